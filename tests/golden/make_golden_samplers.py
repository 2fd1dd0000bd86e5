#!/usr/bin/env python
"""Generate tests/golden/samp_*.npz from the UNMODIFIED reference's pure-PyTorch sampler statements
`grid_scatter_ref` (drtk/grid_scatter.py:108-191) and `mipmap_grid_sample_ref` (drtk/mipmap_grid_sample.py:130-236),
float64 on CPU, values + autograd gradients of a seeded linear loss.

    python tests/golden/make_golden_samplers.py        (build container: /root/reference must exist)

The two files are loaded straight from /root/reference; their module-level `load_torch_ops(...)` (which would load
the CUDA extension) is stubbed out -- only the *_ref functions are executed.  Cases are restricted to where those
statements and the reference's CUDA kernels are documented to coincide (drtk/mipmap_grid_sample.py:143-146: square
textures, force_max_aniso, clip_grad off; grid_scatter bicubic with border/reflection padding only for sample
positions inside the image).  Everything else is pinned on reference-CUDA outputs (make_golden_samplers_cuda.py).
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
pkg = types.ModuleType("drtk"); pkg.__path__ = []
utils = types.ModuleType("drtk.utils"); utils.load_torch_ops = lambda name: None
sys.modules.update({"drtk": pkg, "drtk.utils": utils})


def load(name):
    spec = importlib.util.spec_from_file_location(f"drtk.{name}", f"/root/reference/drtk/{name}.py")
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


GS, MM = load("grid_scatter"), load("mipmap_grid_sample")
dt = th.float64


def scatter_cases():
    i = 0
    for mode in ("bilinear", "bicubic"):
        for pad in ("zeros", "border", "reflection"):
            for align in (False, True):
                i += 1
                g = th.Generator().manual_seed(7000 + i)
                N, C, H, W, Ho, Wo = 2, 3, 9, 11, 8, 10
                x = th.rand((N, C, H, W), generator=g, dtype=dt)
                reach = 1.3 if (mode == "bilinear" or pad == "zeros") else 0.8  # see module docstring
                grid = (th.rand((N, H, W, 2), generator=g, dtype=dt) * 2 - 1) * reach
                w = th.rand((N, C, Ho, Wo), generator=g, dtype=dt)
                xl, gl = x.clone().requires_grad_(True), grid.clone().requires_grad_(True)
                out = GS.grid_scatter_ref(xl, gl, Ho, Wo, mode, pad, align)
                gx, gg = th.autograd.grad((out * w).sum(), (xl, gl))
                yield f"samp_scatter_{mode}_{pad}_{int(align)}", dict(
                    input=x, grid=grid, w=w, out=out.detach(), g_input=gx, g_grid=gg,
                    meta=np.array([Ho, Wo, {"bilinear": 0, "bicubic": 2}[mode], {"zeros": 0, "border": 1, "reflection": 2}[pad], int(align)]))


def mipmap_cases():
    i = 0
    for mode in ("bilinear", "bicubic"):
        for pad in ("zeros", "border", "reflection"):
            for max_aniso, nlev in ((1, 1), (1, 3), (4, 3), (3, 2)):
                i += 1
                g = th.Generator().manual_seed(8000 + i)
                N, C, H, W, S = 2, 3, 7, 9, 16
                levels = [th.rand((N, C, S >> l, S >> l), generator=g, dtype=dt) for l in range(nlev)]
                grid = (th.rand((N, H, W, 2), generator=g, dtype=dt) * 2 - 1) * 1.1
                # footprints from well below one texel to several texels, anisotropic
                jac = (th.rand((N, H, W, 2, 2), generator=g, dtype=dt) - 0.5) * th.tensor([0.5, 0.05], dtype=dt)[:, None]
                jac = jac * th.rand((N, H, W, 1, 1), generator=g, dtype=dt)
                w = th.rand((N, C, H, W), generator=g, dtype=dt)
                ll = [t.clone().requires_grad_(True) for t in levels]
                gl = grid.clone().requires_grad_(True)
                out = MM.mipmap_grid_sample_ref(ll, gl, jac, max_aniso, mode, pad, False)
                grads = th.autograd.grad((out * w).sum(), ll + [gl], allow_unused=True)
                d = dict(grid=grid, jac=jac, w=w, out=out.detach(), g_grid=grads[-1],
                         meta=np.array([max_aniso, {"bilinear": 0, "bicubic": 2}[mode], {"zeros": 0, "border": 1, "reflection": 2}[pad], nlev]))
                for l in range(nlev):
                    d[f"level{l}"] = levels[l]
                    d[f"g_level{l}"] = grads[l] if grads[l] is not None else th.zeros_like(levels[l])
                yield f"samp_mipmap_{mode}_{pad}_a{max_aniso}_l{nlev}", d


for name, d in list(scatter_cases()) + list(mipmap_cases()):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: (v.detach().numpy() if th.is_tensor(v) else v) for k, v in d.items()})
    print(name)
