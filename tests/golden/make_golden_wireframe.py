"""Writes tests/golden/wire_*.npz: inputs + outputs of the REFERENCE CUDA wireframe rasteriser
(rasterize_lines_kernel, src/rasterize/rasterize_kernel.cu:261-400).  The reference has no CPU twin for
wireframe mode (rasterize_kernel_cpu.cpp:257), so this script must run on a GPU box that carries
oracle/_ref/*.so:   gpurun -- 'python tests/golden/make_golden_wireframe.py gpurun_out/golden'
and the files are then copied into tests/golden/.  Nothing of drtk_b200's kernels is involved."""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from drtk_b200 import scenes  # noqa: E402  (pure-torch scene generators only)
from oracle import ref as R  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden")
os.makedirs(out, exist_ok=True)
assert R.available() and th.cuda.is_available()


def flags(vi, seed):
    g = th.Generator().manual_seed(seed or 0)
    f = th.randint(0, 8, (vi.shape[0],), generator=g, dtype=th.int64)
    if seed is None:  # all three edges visible
        f = th.full_like(f, 7)
    o = vi.clone().to(th.int64)
    o[:, 0] = o[:, 0] | (f << 28)
    return o.to(th.int32)


cases = {
    "wire_grid_96": (*scenes.grid_mesh(9, 9, 96, 96, 2, seed=31), 96, 96, 3),
    "wire_overdraw_80x64": (*scenes.grid_mesh(7, 6, 64, 80, 2, seed=37, overdraw=True), 64, 80, 4),
    "wire_two_tri_128": (scenes.two_triangles()[0] / 4.0, scenes.two_triangles()[1], 128, 128, None),
}
for name, (v, vi, H, W, seed) in cases.items():
    if name == "wire_two_tri_128":
        v = v.clone(); v[..., 2] = v[..., 2] * 4.0
    vif = flags(vi, seed)  # the nibble goes into column 0 only
    d, i = R.rasterize_with_depth(v.cuda(), vif.cuda(), H, W, wireframe=True)
    np.savez_compressed(os.path.join(out, name + ".npz"), v=v.numpy(), vi=vif.numpy(), H=H, W=W,
                        index_img=i.cpu().numpy(), depth_img=d.cpu().numpy())
    print(name, "line px", int((i >= 0).sum()), "occluder px", int(((i < 0) & (d > 0)).sum()))
