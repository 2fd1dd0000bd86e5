"""screen_space_uv_derivative (producer of mipmap_grid_sample's `vt_dxdy_img`).

Fixtures tests/golden/uvd_*.npz: outputs of the UNMODIFIED reference function on CPU (its Python files loaded from
/root/reference, its `interpolate` = the reference's CPU kernel; tests/golden/make_golden_uv_derivative.py), fp32.
CPU suite: numpy oracle (float64) vs fixtures.  GPU suite: the fused CUDA kernel vs fixtures / oracle, and vs the
differentiable composition built on this package's interpolate (which is also what a gradient request runs).
"""
import glob
import os

import numpy as np
import pytest
import torch as th

import drtk_b200
from oracle import oracle as O
from tests.util import GOLDEN, assert_close

NAMES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "uvd_*.npz")))
KEYS = ("v", "vt", "vi", "vti", "index_img", "bary_img", "mask", "campos", "camrot", "focal")
DEV = "cuda:0"


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


@pytest.mark.parametrize("name", NAMES)
def test_oracle_matches_reference(name):
    g = load(name)
    assert len(NAMES) >= 2
    out = O.screen_space_uv_derivative(*[g[k] for k in KEYS])
    assert_close(out, g["out"], rtol=2e-4, what="vt_dxdy_img")  # the fixture is a float32 chain of ~12 ops incl. two inverses


def test_face_dpdt_and_cpu_composition():
    g = load(NAMES[0])
    t = {k: th.from_numpy(g[k]) for k in KEYS}
    dpdt, corners = drtk_b200.utils.face_dpdt(t["v"].double(), t["vt"].double(), t["vi"].long(), t["vti"].long())
    # defining property: (dp/dt)^T maps uv edges back onto position edges
    uv = t["vt"].double()[:, t["vti"].long()]
    e_t, e_p = uv[:, :, 1:3] - uv[:, :, 0:1], corners[:, :, 1:3] - corners[:, :, 0:1]
    assert th.allclose(e_t @ dpdt, e_p, atol=1e-9)
    with pytest.raises(RuntimeError):  # the composition needs interpolate, which has no CPU path
        drtk_b200.screen_space_uv_derivative(*[t[k] for k in KEYS])


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_fused_matches_reference_and_composition(name):
    g = load(name)
    t = {k: th.from_numpy(g[k]).to(DEV) for k in KEYS}
    args = [t[k] for k in KEYS]
    fused = drtk_b200.screen_space_uv_derivative(*args)
    assert fused.shape == g["out"].shape and not fused.requires_grad
    assert_close(fused.cpu().numpy(), g["out"], rtol=2e-4, what="fused vs reference")
    # float32 through two 2x2 inverses: a few ill-conditioned uv triangles lose digits (the reference's own fp32
    # chain sits as far from the float64 oracle); the bulk is at fp32 round-off
    exact = O.screen_space_uv_derivative(*[g[k] for k in KEYS])
    assert_close(fused.cpu().numpy(), exact, rtol=3e-4, what="fused vs oracle")
    err = np.abs(fused.cpu().numpy() - exact)
    assert (err > 2e-5 * np.abs(exact) + 2e-5 * np.abs(exact).max()).mean() < 0.01
    # a gradient request takes the differentiable composition; same values, and the gradient reaches v / vt
    t["v"].requires_grad_(True); t["vt"].requires_grad_(True)
    comp = drtk_b200.screen_space_uv_derivative(*[t[k] for k in KEYS])
    assert comp.requires_grad
    assert_close(comp.detach().cpu().numpy(), fused.cpu().numpy(), rtol=2e-4, what="composition vs fused")
    comp.sum().backward()
    assert float(t["v"].grad.abs().sum()) > 0 and float(t["vt"].grad.abs().sum()) > 0
    # strided bary / int64 indices / a partial mask
    bary_nc = t["bary_img"].permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    mask2 = t["mask"].clone(); mask2[:, ::2] = False
    with th.no_grad():
        out2 = drtk_b200.screen_space_uv_derivative(t["v"], t["vt"], t["vi"].long(), t["vti"].long(), t["index_img"], bary_nc,
                                                    mask2, t["campos"], t["camrot"], t["focal"])
    assert th.equal(out2[:, 1::2], fused[:, 1::2]) and float(out2[:, ::2].abs().max()) == 0.0


@pytest.mark.gpu
def test_cuda_uv_derivative_feeds_mipmap_grid_sample():
    """The chain the function exists for: rasterize -> render -> uv interpolate -> uv derivative -> mipmap lookup."""
    g = load(NAMES[0])
    t = {k: th.from_numpy(g[k]).to(DEV) for k in KEYS}
    vt_img = drtk_b200.interpolate(t["vt"][:, :t["v"].shape[1]].contiguous(), t["vi"], t["index_img"], t["bary_img"])
    grid = vt_img.permute(0, 2, 3, 1) * 2 - 1
    jac = drtk_b200.screen_space_uv_derivative(*[t[k] for k in KEYS])
    tex = [th.rand(2, 3, 64 >> l, 64 >> l, device=DEV) for l in range(4)]
    out = drtk_b200.mipmap_grid_sample(tex, grid, jac, 4, padding_mode="border")
    assert out.shape == (2, 3) + tuple(t["index_img"].shape[1:]) and bool(th.isfinite(out).all())
