#!/usr/bin/env python
"""Writes tests/golden/sampcuda_*.npz: inputs + outputs (forward and backward) of the REFERENCE CUDA sampler kernels
(mipmap_grid_sampler_ext / grid_scatter_ext built from the unmodified reference into oracle/_ref) for the behaviour
the reference's torch statements cannot express: per-pixel anisotropic sample counts, clip_grad, the forward's
align_corners override, non-square textures, grid_scatter's bicubic padding of out-of-image sample positions.
The reference has no CPU twin for these ops, so this runs on a GPU box:
    gpurun -- 'python tests/golden/make_golden_samplers_cuda.py gpurun_out/golden'
and the files are then copied into tests/golden/.  Nothing of drtk_b200's kernels is involved."""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as R  # noqa: E402

out_dir = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out_dir, exist_ok=True)
assert R.samplers_available() and th.cuda.is_available()
dev = "cuda:0"
MODE = {"bilinear": 0, "bicubic": 2}
PAD = {"zeros": 0, "border": 1, "reflection": 2}

# (name, mode, pad, max_aniso, align, force, clip, levels)
MIP = [("aniso", "bilinear", "zeros", 6, False, False, False, 3),
       ("clip", "bilinear", "border", 4, False, False, True, 2),
       ("align", "bilinear", "reflection", 3, True, False, False, 3),
       ("bicubic_clip", "bicubic", "border", 4, False, False, True, 2),
       ("bicubic_aniso", "bicubic", "zeros", 5, True, False, False, 3)]
for i, (name, mode, pad, aniso, align, force, clip, nlev) in enumerate(MIP):
    g = th.Generator().manual_seed(9100 + i)
    N, C, H, W, SH, SW = 1, 2, 24, 28, 32, 24
    levels = [th.rand((N, C, SH >> l, SW >> l), generator=g) for l in range(nlev)]
    grid = (th.rand((N, H, W, 2), generator=g) * 2 - 1) * 1.15
    jac = (th.rand((N, H, W, 2, 2), generator=g) - 0.5) * th.tensor([0.6, 0.06])[:, None] * th.rand((N, H, W, 1, 1), generator=g)
    w = th.rand((N, C, H, W), generator=g)
    lv = [t.to(dev).requires_grad_(True) for t in levels]
    gr = grid.to(dev).requires_grad_(True)
    out = R.mipmap_grid_sample(lv, gr, jac.to(dev), aniso, mode, pad, align, force, clip)
    (out * w.to(dev)).sum().backward()
    d = dict(grid=grid, jac=jac, w=w, out=out.detach().cpu(), g_grid=gr.grad.cpu(),
             meta=np.array([aniso, MODE[mode], PAD[pad], nlev, int(align), int(force), int(clip)]))
    for l in range(nlev):
        d[f"level{l}"] = levels[l]
        d[f"g_level{l}"] = lv[l].grad.cpu()
    np.savez_compressed(os.path.join(out_dir, f"sampcuda_mipmap_{name}.npz"), **{k: np.asarray(v) for k, v in d.items()})
    print("mipmap", name)

for i, (mode, pad, align) in enumerate([("bicubic", "border", False), ("bicubic", "reflection", False),
                                        ("bicubic", "border", True), ("bilinear", "reflection", True)]):
    g = th.Generator().manual_seed(9200 + i)
    N, C, H, W, Ho, Wo = 1, 2, 18, 22, 13, 17   # prime-ish sizes: no exact flip boundaries in the reflection
    x = th.rand((N, C, H, W), generator=g)
    grid = (th.rand((N, H, W, 2), generator=g) * 2 - 1) * 1.6
    w = th.rand((N, C, Ho, Wo), generator=g)
    xl, gl = x.to(dev).requires_grad_(True), grid.to(dev).requires_grad_(True)
    out = R.grid_scatter(xl, gl, Ho, Wo, mode, pad, align)
    (out * w.to(dev)).sum().backward()
    np.savez_compressed(os.path.join(out_dir, f"sampcuda_scatter_{mode}_{pad}_{int(align)}.npz"), input=x.numpy(),
                        grid=grid.numpy(), w=w.numpy(), out=out.detach().cpu().numpy(), g_input=xl.grad.cpu().numpy(),
                        g_grid=gl.grad.cpu().numpy(), meta=np.array([Ho, Wo, MODE[mode], PAD[pad], int(align)]))
    print("scatter", mode, pad, align)
