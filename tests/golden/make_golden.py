#!/usr/bin/env python
"""Generate tests/golden/*.npz and known_answers.json from the UNMODIFIED reference.

Run in the build container (where /root/reference exists) after `python oracle/build_ref.py`:

    python tests/golden/make_golden.py

The reference's own kernels (CPU twins inside oracle/_ref/*.so, compiled from the reference
sources where they lie) are executed on small deterministic scenes; inputs and outputs are
stored so that the tests can run where the reference is absent (the GPU box).  The reference
repository ships no golden vectors of its own; these fixtures, produced by the
reference itself, are what pins the oracle and the CUDA kernels.
"""
import json
import os
import sys
import zlib

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from drtk_b200 import scenes  # noqa: E402
from oracle import ref as R  # noqa: E402


def scene_defs():
    v2, vi2, _, _ = scenes.two_triangles()
    v2 = v2 / 4.0  # dyadic scaling keeps every edge function exact: 128 x 128 canvas
    yield "two_tri_128", v2, vi2, 128, 128
    v, vi = scenes.grid_mesh(9, 9, 48, 48, 2, seed=101)
    yield "grid_48", v, vi, 48, 48
    v, vi = scenes.grid_mesh(7, 7, 48, 64, 1, seed=202, overdraw=True)
    yield "overdraw_64x48", v, vi, 48, 64
    # integer-coordinate fan: pixel centres exactly on shared edges / vertices (top-left rule,
    # watertightness, lowest-id tie break on a duplicated triangle)
    v = th.tensor([[[8, 8, 1], [24, 8, 1], [24, 24, 1], [8, 24, 1], [16, 16, 1], [16, 16, 2]]], dtype=th.float32)
    vi = th.tensor([[0, 1, 4], [1, 2, 4], [2, 3, 4], [3, 0, 4], [0, 1, 4], [0, 2, 5], [3, 3, 3]], dtype=th.int32)
    yield "fan_32", v, vi, 32, 32


def run_reference(v, vi, H, W, C=3, seed=5):
    N, V = v.shape[:2]
    depth_r, index = R.rasterize_with_depth(v, vi, H, W)
    attr = scenes.vertex_attributes(N, V, C, seed=seed)
    gen = th.Generator().manual_seed(seed + 1)
    w_img = th.rand((N, C, H, W), generator=gen)
    w_depth = th.rand((N, H, W), generator=gen)
    w_bary = th.rand((N, 3, H, W), generator=gen)
    vv = v.clone().requires_grad_(True)
    aa = attr.clone().requires_grad_(True)
    depth, bary = R.render(vv, vi, index)
    img = R.interpolate(aa, vi, index, bary)
    cap = {}
    img2 = R.edge_grad_estimator(vv, vi, bary, img, index, v_pix_img_hook=lambda g: cap.__setitem__("g", g.clone()))
    # individual op gradients (each op's backward in isolation)
    ga, gb = th.autograd.grad((img * w_img).sum(), (aa, bary), retain_graph=True)
    gv_render, = th.autograd.grad((bary * w_bary).sum() + (depth * w_depth).sum(), vv, retain_graph=True)
    # full pipeline gradient
    loss = (img2 * w_img).sum()
    gv_full, ga_full = th.autograd.grad(loss, (vv, aa))
    return dict(
        v=v, vi=vi, attr=attr, w_img=w_img, w_depth=w_depth, w_bary=w_bary,
        raster_depth=depth_r, index_img=index, depth_img=depth.detach(), bary_img=bary.detach(),
        interp=img.detach(), grad_attr=ga, grad_bary=gb, grad_v_render=gv_render,
        grad_v_pix_img=cap["g"], grad_v_full=gv_full, grad_attr_full=ga_full,
    )


def main():
    th.set_num_threads(1)  # fixed accumulation order inside the reference CPU twins
    for name, v, vi, H, W in scene_defs():
        out = run_reference(v, vi, H, W)
        arrs = {k: t.numpy() for k, t in out.items()}
        arrs["HW"] = np.array([H, W], np.int64)
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **arrs)
        print(f"{name}: covered {int((arrs['index_img'] >= 0).sum())} px, {os.path.getsize(path) / 1e3:.0f} kB")

    # full-size known answers (scalars only)
    ka = {}
    for name, (v, vi, H, W) in {"hello_triangle_512": scenes.hello_triangle(), "two_triangles_512": scenes.two_triangles()}.items():
        d, idx = R.rasterize_with_depth(v, vi, H, W)
        idx_np = idx.numpy()
        ka[name] = dict(
            covered=int((idx_np >= 0).sum()),
            per_triangle=[int((idx_np == t).sum()) for t in range(vi.shape[0])],
            index_crc32=zlib.crc32(idx_np.tobytes()),
            depth_sum=float(d.double().sum()),
        )
    v, vi, H, W = scenes.config_mesh(3, N=1)
    d, idx = R.rasterize_with_depth(v, vi, H, W)
    ka["config3_n1_1024"] = dict(covered=int((idx.numpy() >= 0).sum()), index_crc32=zlib.crc32(idx.numpy().tobytes()))
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f, indent=1, sort_keys=True)
    print(json.dumps(ka, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
