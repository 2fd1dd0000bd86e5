"""float64 dispatch (csrc/fp64.cu): the reference instantiates every hot-path kernel for double
(src/include/kernel_utils.h:47-57).  GPU suite: the double kernels through the public API / C ABI against the f64
build of the CPU oracle (pinned on the reference's CPU twins, tests/test_oracle_golden.py) to 1e-10, against the
reference's own CUDA double kernels when oracle/_ref travelled, and against the fp32 path (consistency)."""
import numpy as np
import pytest
import torch as th

import drtk_b200
from drtk_b200 import scenes
from oracle import oracle as O
from oracle import ref as R
from tests.util import assert_close

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/*.so (reference build) not present")

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def scene(name):
    if name == "grid":
        v, vi = scenes.grid_mesh(17, 13, 120, 150, 2, seed=3)
        return v, vi, 120, 150
    if name == "overdraw":
        v, vi = scenes.grid_mesh(11, 11, 96, 80, 2, seed=5, overdraw=True)
        return v, vi, 96, 80
    v, vi, H, W = scenes.two_triangles()
    return v / 4.0, vi, H // 4, W // 4


def run_pipeline(api, v, vi, attr, w, H, W, dt, hook):
    vv = v.to(DEV, dt).requires_grad_(True)
    aa = attr.to(DEV, dt).requires_grad_(True)
    vid = vi.to(DEV)
    depth_r, index = api.rasterize_with_depth(vv, vid, H, W)
    depth, bary = api.render(vv, vid, index)
    img = api.interpolate(aa, vid, index, bary)
    cap = {}
    out = api.edge_grad_estimator(vv, vid, bary, img, index, v_pix_img_hook=(lambda g: cap.__setitem__("g", g.clone())) if hook else None)
    ((out * w.to(DEV, dt)).sum() + depth.sum()).backward()
    r = dict(depth_r=depth_r, index=index, depth=depth, bary=bary, img=img, gv=vv.grad, ga=aa.grad, gimg=cap.get("g"))
    return {k: (t.detach() if t is not None else None) for k, t in r.items()}


@pytest.mark.parametrize("name", ["grid", "overdraw", "two_tri"])
@pytest.mark.parametrize("hook", [False, True])
def test_f64_pipeline_vs_oracle(name, hook):
    v, vi, H, W = scene(name)
    N, V, C = v.shape[0], v.shape[1], 5
    attr = scenes.vertex_attributes(N, V, C, seed=9)
    w = th.rand((N, C, H, W), generator=th.Generator().manual_seed(4))
    r = run_pipeline(drtk_b200, v, vi, attr, w, H, W, th.float64, hook)
    assert r["depth_r"].dtype == th.float32 and r["index"].dtype == th.int32  # :481
    for k in ("depth", "bary", "img", "gv", "ga"):
        assert r[k].dtype == th.float64
    v64, a64, vin = v.double().numpy(), attr.double().numpy(), vi.numpy()
    d_o, idx_o = O.rasterize(v64, vin, H, W, mode=0)
    idx = r["index"].cpu().numpy()
    assert (idx != idx_o).mean() < 2e-4  # fused vs unfused products can flip a sample lying exactly on an edge
    assert_close(r["depth_r"].cpu().numpy()[idx == idx_o], d_o[idx == idx_o], rtol=1e-6, what="raster depth")
    d, b = O.render_fwd(v64, vin, idx)
    assert_close(r["depth"].cpu().numpy(), d, rtol=1e-11, what="depth")
    assert_close(r["bary"].cpu().numpy(), b, rtol=1e-11, what="bary")
    img = O.interpolate_fwd(a64, vin, idx, b)
    assert_close(r["img"].cpu().numpy(), img, rtol=1e-11, what="img")
    gpix = O.edge_grad_bwd(v64, img, idx, vin, w.double().numpy(), 1e4)
    if hook:
        assert_close(r["gimg"].cpu().numpy(), gpix, rtol=1e-9, what="grad_v_pix_img")
    gv_e, _ = O.interpolate_bwd(gpix, v64, vin, idx, b, True, False)
    ga, gb = O.interpolate_bwd(w.double().numpy(), a64, vin, idx, b, True, True)
    gv = gv_e + O.render_bwd(v64, vin, idx, np.ones_like(d), gb)
    assert_close(r["ga"].cpu().numpy(), ga, rtol=1e-10, what="grad attr")
    assert_close(r["gv"].cpu().numpy(), gv, rtol=1e-8, what="grad v")


def test_f64_agrees_with_f32_path():
    v, vi, H, W = scene("grid")
    attr = scenes.vertex_attributes(v.shape[0], v.shape[1], 4, seed=2)
    w = th.rand((v.shape[0], 4, H, W), generator=th.Generator().manual_seed(1))
    a = run_pipeline(drtk_b200, v, vi, attr, w, H, W, th.float64, False)
    b = run_pipeline(drtk_b200, v, vi, attr, w, H, W, th.float32, False)
    assert (a["index"] != b["index"]).float().mean() < 2e-4
    same = (a["index"] == b["index"]).cpu().numpy()
    assert_close(b["img"].cpu().numpy()[:, :, same[0] & same[1]], a["img"].cpu().numpy()[:, :, same[0] & same[1]], rtol=2e-5)
    assert_close(b["ga"].cpu().numpy(), a["ga"].cpu().numpy(), rtol=1e-4, what="grad attr f32 vs f64")


@pytest.mark.skipif(not R.available(), reason="oracle/_ref not present")
def test_f64_vs_reference_cuda_f64():
    v, vi, H, W = scene("overdraw")
    attr = scenes.vertex_attributes(v.shape[0], v.shape[1], 3, seed=7)
    w = th.rand((v.shape[0], 3, H, W), generator=th.Generator().manual_seed(8))
    a = run_pipeline(drtk_b200, v, vi, attr, w, H, W, th.float64, True)
    b = run_pipeline(R, v, vi, attr, w, H, W, th.float64, True)
    assert (a["index"] != b["index"]).float().mean() < 2e-4
    if th.equal(a["index"], b["index"]):
        # background pixels of `img` carry the float coordinate sweep (interpolate_kernel.cu:104-109), whose division
        # the reference build approximates (--use_fast_math): one float ulp there, exact elsewhere
        fg = (a["index"] >= 0)[:, None].expand_as(a["img"]).cpu().numpy()
        assert_close(a["img"].cpu().numpy()[fg], b["img"].cpu().numpy()[fg], rtol=1e-10, what="img (covered pixels)")
        assert_close(a["img"].cpu().numpy()[~fg], b["img"].cpu().numpy()[~fg], rtol=3e-7, what="img (background sweep)")
        for k, tol in (("depth", 1e-10), ("bary", 1e-10), ("gimg", 1e-7), ("ga", 1e-9), ("gv", 1e-6)):
            assert_close(a[k].cpu().numpy(), b[k].cpu().numpy(), rtol=tol, what=k)


def test_f64_strided_and_errors():
    v, vi, H, W = scene("two_tri")
    vb = th.zeros(v.shape[0], v.shape[1], 4, dtype=th.float64, device=DEV)
    vb[..., :3] = v.to(DEV)
    vs = vb[..., :3]
    vid = vi.to(DEV)
    i1 = drtk_b200.rasterize(vs, vid, H, W)
    i2 = drtk_b200.rasterize(vs.contiguous(), vid, H, W)
    assert th.equal(i1, i2)
    d1, b1 = drtk_b200.render(vs, vid, i1)
    d2, b2 = drtk_b200.render(vs.contiguous(), vid, i1)
    assert th.equal(b1, b2) and th.equal(d1, d2)
    # wireframe is served in double as well (the reference instantiates rasterize_lines_kernel<double>)
    iw1 = drtk_b200.rasterize(vs, vid, H, W, wireframe=True)
    iw2 = drtk_b200.rasterize(vs.contiguous(), vid, H, W, wireframe=True)
    assert th.equal(iw1, iw2)
    with pytest.raises(RuntimeError, match="same dtype"):
        drtk_b200.interpolate(vs.float(), vid, i1, b1)


@needs_ref
@pytest.mark.parametrize("name", ["grid", "overdraw", "two_tri"])
def test_f64_wireframe_vs_reference_cuda_and_f32(name):
    """rasterize(..., wireframe=True) for float64 vertices (src/rasterize/rasterize_kernel.cu:492-535 dispatches
    rasterize_lines_kernel<double>): index_img / depth_img against the reference's CUDA double kernel, and -- away from
    knife-edge diamond crossings -- against this package's bit-exact float32 wireframe."""
    v, vi, H, W = scene(name)
    vid = vi.to(DEV).clone()
    vid[:, 0] |= (7 << 28)  # all three edges visible (:293-303)
    v64 = v.to(DEV).double()
    d, i = drtk_b200.rasterize_with_depth(v64, vid, H, W, wireframe=True)
    dr, ir = R.rasterize_with_depth(v64, vid, H, W, wireframe=True)
    assert d.dtype == th.float32 and i.dtype == th.int32
    assert int((i >= 0).sum()) > 0
    # double arithmetic, same expressions: the discrete decisions agree except (at most) a handful of exact ties
    assert int((i != ir).sum()) <= max(2, int(1e-5 * i.numel())), f"{int((i != ir).sum())} pixels differ from the reference"
    same = i == ir
    assert float((d - dr).abs()[same].max()) <= 1e-6 * float(dr.abs().max())
    d32, i32 = drtk_b200.rasterize_with_depth(v.to(DEV), vid, H, W, wireframe=True)
    assert float((i != i32).float().mean()) < 2e-3
