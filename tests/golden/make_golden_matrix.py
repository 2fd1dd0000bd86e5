"""Writes tests/golden/mat_*.npz: inputs + outputs of the REFERENCE CPU implementations of
interpolation_matrix / interpolation_normal_matrix (oracle/_ref/interpolate_ext.so, built from the unmodified
reference sources; ops `interpolate_ext::interpolation_matrix`, `interpolate_ext::interpolation_normal_matrix`,
src/interpolate/interpolate_module.cpp:635-640), including the gradients w.r.t. bary_img of a random linear
functional of the values.  Run where /root/reference was built:  python tests/golden/make_golden_matrix.py"""
import os
import sys

import numpy as np
import torch as th

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from drtk_b200 import scenes  # noqa: E402
from oracle import ref as R  # noqa: E402

R.load()
out = os.path.join(ROOT, "tests", "golden")
cases = {
    "mat_grid_40x56": (*scenes.grid_mesh(7, 6, 40, 56, 2, seed=51, overdraw=True), 40, 56),
    "mat_two_tri_64": (scenes.two_triangles()[0] / 8.0, scenes.two_triangles()[1], 64, 64),
}
for name, (v, vi, H, W) in cases.items():
    if name == "mat_two_tri_64":
        v = v.clone(); v[..., 2] *= 8.0
    N, V = v.shape[0], v.shape[1]
    vin = vi[None].expand(N, -1, -1).contiguous()
    index = R.rasterize(v, vin, H, W)
    _, bary = R.render(v, vin, index)
    g = th.Generator().manual_seed(7)
    b1 = bary.clone().requires_grad_(True)
    crow, col, val, rows = th.ops.interpolate_ext.interpolation_matrix(vin, index, b1)
    w1 = th.rand(val.shape, generator=g)
    (val * w1).sum().backward()
    b2 = bary.clone().requires_grad_(True)
    ncrow, ncol, nval = th.ops.interpolate_ext.interpolation_normal_matrix(vin, index, b2, V)
    w2 = th.rand(nval.shape, generator=g)
    (nval * w2).sum().backward()
    np.savez_compressed(os.path.join(out, name + ".npz"), vi=vi.numpy(), index_img=index.numpy(), bary_img=bary.numpy(), V=V,
                        crow=crow.numpy(), col=col.numpy(), values=val.detach().numpy(), row_pixels=rows.numpy(),
                        w_values=w1.numpy(), grad_bary=b1.grad.numpy(),
                        n_crow=ncrow.numpy(), n_col=ncol.numpy(), n_values=nval.detach().numpy(),
                        n_w_values=w2.numpy(), n_grad_bary=b2.grad.numpy())
    print(name, "rows", rows.numel(), "nnz(normal)", nval.numel())
