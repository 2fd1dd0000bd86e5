"""GPU suite (-m gpu): the CUDA kernels, called through the C ABI (drtk_b200/_ops.py -> ctypes ->
libdrtk_b200.so), against
  (1) the golden vectors produced by the reference itself (tests/golden/*.npz),
  (2) the CPU oracle (oracle/), on seeded inputs it finishes in seconds,
  (3) the reference's own CUDA kernels (oracle/_ref/*.so, built from the unmodified reference
      sources) when they travelled to the box: index_img / depth_img BIT-EXACT,
  (4) size-independent properties at BASELINE.json sizes.
Tolerances: integer outputs bit-exact; fp32 outputs rtol 1e-5 (north star) with the same fraction
of the tensor's scale as absolute floor (tests/util.py:assert_close).
"""
import json
import os
import zlib

import numpy as np
import pytest
import torch as th

import drtk_b200
from drtk_b200 import _ops, scenes
from oracle import oracle as O
from oracle import ref as R
from tests.util import GOLDEN, assert_close, golden_cases, load_golden, ulp_diff

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CASES = golden_cases()
HAVE_REF = R.available()
needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/*.so (reference build) not present")


def cu(x, dtype=None):
    t = th.as_tensor(x)
    if dtype is not None:
        t = t.to(dtype)
    return t.to(DEV)


def npy(t):
    return t.detach().cpu().numpy()


def assert_close_device(actual, expected, rtol, what):
    """assert_close on the device (tensors too large to ship to numpy): |a-e| <= rtol*|e| + rtol*max|e|, written
    as NOT(within tolerance) so that a NaN / Inf in `actual` fails instead of slipping through a `>` test."""
    assert actual.shape == expected.shape, f"{what}: shape {tuple(actual.shape)} vs {tuple(expected.shape)}"
    assert bool(th.isfinite(expected).all()), f"{what}: the expectation itself is not finite"
    scale = float(expected.abs().max()) if expected.numel() else 0.0
    err = (actual - expected).abs()
    bad = ~(err <= rtol * expected.abs() + rtol * scale)
    if bool(bad.any()):
        i = int(th.where(bad.reshape(-1), th.nan_to_num(err.reshape(-1), nan=float("inf")), th.zeros((), device=err.device)).argmax())
        raise AssertionError(f"{what}: {int(bad.sum())}/{bad.numel()} elements out of tolerance (or not finite); worst: actual "
                             f"{float(actual.reshape(-1)[i])!r} expected {float(expected.reshape(-1)[i])!r}, scale {scale:.3e}")


def assert_edge_gradient_image_close(actual, expected, what):
    """grad_v_pix_img [N,3,H,W] against the reference kernel's.  Pixel pairs on an INTERSECTION of two surfaces go through
    get_dp_dr (src/edge_grad/edge_grad_kernel.cu:102-203): a division by the sine of the angle between two face normals
    (clamped at max_dp_dr = 1e4), computed with MUFU rsqrt / rcp under --use_fast_math.  Where the two normals are nearly
    parallel that quotient is ill-conditioned in BOTH implementations: a last-ulp difference in a normal moves the result by
    up to ~1e-2 of itself (measured at config 4 overdraw-2: 240 of 1e8 elements beyond 1e-5, worst 2.2e-3 relative).
    Rule: every element within 1e-5 (relative + 1e-5 of the typical non-zero magnitude), except at most one element in
    100 000, which must still agree to 2e-2 of its own size.  NaN / Inf fail."""
    assert actual.shape == expected.shape and bool(th.isfinite(expected).all())
    nz = expected[expected != 0].abs()
    typical = float(nz.median()) if nz.numel() else 0.0
    err = (actual - expected).abs()
    tight = err <= 1e-5 * expected.abs() + 1e-5 * typical
    loose = err <= 2e-2 * expected.abs() + 1e-5 * typical
    n_beyond = int((~tight).sum())
    if n_beyond > max(1, expected.numel() // 100_000) or not bool(loose.all()):
        bad = ~loose if not bool(loose.all()) else ~tight
        i = int(th.where(bad.reshape(-1), th.nan_to_num(err.reshape(-1), nan=float("inf")), th.zeros((), device=err.device)).argmax())
        raise AssertionError(f"{what}: {n_beyond}/{expected.numel()} elements beyond 1e-5, {int((~loose).sum())} beyond 2e-2; worst: "
                             f"actual {float(actual.reshape(-1)[i])!r} expected {float(expected.reshape(-1)[i])!r}, typical {typical:.3e}")


def test_comparison_helpers_reject_nan_and_inf():
    """Self-test of the parity tooling: a kernel that emits NaN / Inf must FAIL every fp32 comparison."""
    e = np.array([1.0, 1.0, 2.0, 0.0], np.float32)
    for bad_val in (np.nan, np.inf, -np.inf):
        a = e.copy(); a[2] = bad_val
        with pytest.raises(AssertionError):
            assert_close(a, e, what="nan-injection")
        with pytest.raises(AssertionError):
            assert_close_device(cu(a), cu(e), 1e-5, "nan-injection (device)")
    assert_close(e, e)
    assert_close_device(cu(e), cu(e), 1e-5, "identity")
    # a NaN injected into a real kernel output is caught too
    v, vi = scenes.grid_mesh(9, 9, 64, 64, 1, seed=3)
    index = drtk_b200.rasterize(cu(v), cu(vi), 64, 64)
    _, bary = drtk_b200.render(cu(v), cu(vi), index)
    poisoned = bary.clone(); poisoned[0, 1, 30, 30] = float("nan")
    with pytest.raises(AssertionError):
        assert_close(npy(poisoned), npy(bary), what="poisoned bary")


def small_scenes():
    yield "grid_256", *scenes.grid_mesh(21, 21, 256, 256, 2, seed=7), 256, 256
    yield "overdraw_192x160", *scenes.grid_mesh(15, 15, 192, 160, 2, seed=11, overdraw=True), 192, 160
    yield "odd_size_97x61", *scenes.grid_mesh(9, 7, 61, 97, 3, seed=13, overdraw=True), 61, 97
    yield "tiny_3x5", *scenes.grid_mesh(3, 3, 3, 5, 1, seed=17), 3, 5
    # a few large, overlapping, partly off-screen triangles (the "large triangle" path of the tiler)
    g = th.Generator().manual_seed(23)
    v = th.rand((2, 30, 3), generator=g) * th.tensor([400.0, 300.0, 3.0]) + th.tensor([-60.0, -40.0, 0.5])
    vi = th.randint(0, 30, (24, 3), generator=g, dtype=th.int32)
    yield "big_tris_280x200", v, vi, 200, 280


SMALL = list(small_scenes())
SMALL_IDS = [s[0] for s in SMALL]


def deep_scene():
    """300 big triangles over one another: every tile matches more boxes than one pass of the tiler holds (the overflow /
    redo path of the box-list scan, several passes per tile, deep z-order).  Rasterize tests only."""
    g = th.Generator().manual_seed(29)
    v = th.rand((1, 900, 3), generator=g) * th.tensor([360.0, 360.0, 3.0]) + th.tensor([-50.0, -50.0, 0.5])
    vi = th.arange(900, dtype=th.int32).view(300, 3)
    return "deep_300_tris_256", v, vi, 256, 256


RASTER = SMALL + [deep_scene()]
RASTER_IDS = [s[0] for s in RASTER]


# ------------------------------------------------------------------------------------------------
# rasterize
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("algo", [0, 1])
def test_rasterize_golden(name, algo):
    g = load_golden(name)
    H, W = map(int, g["HW"])
    depth, index = _ops.rasterize(cu(g["v"]), cu(g["vi"])[None].expand(g["v"].shape[0], -1, -1), H, W, algo=algo)
    np.testing.assert_array_equal(npy(index), g["index_img"])
    assert ulp_diff(npy(depth), g["raster_depth"]).max() <= 8
    assert (npy(depth)[g["index_img"] < 0] == 0).all()


@pytest.mark.parametrize("scene", SMALL, ids=SMALL_IDS)
def test_rasterize_vs_oracle(scene):
    _, v, vi, H, W = scene
    d_o, i_o, margin = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=1, with_margin=True)
    for algo in (0, 1):
        depth, index = _ops.rasterize(cu(v), cu(vi)[None].expand(v.shape[0], -1, -1), H, W, algo=algo)
        mism = npy(index) != i_o
        # MUFU.RCP cannot be reproduced on a CPU: a z-test may legitimately flip only where the two
        # nearest candidates are within a few ulp of each other
        assert not (mism & (margin > 64)).any(), f"algo {algo}: {int((mism & (margin > 64)).sum())} wrong pixels"
        assert mism.sum() <= max(2, 1e-4 * mism.size)
        assert ulp_diff(npy(depth)[~mism], d_o[~mism]).max() <= 16


@pytest.mark.parametrize("scene", RASTER, ids=RASTER_IDS)
def test_rasterize_algorithms_agree_bitwise(scene):
    _, v, vi, H, W = scene
    vi_b = cu(vi)[None].expand(v.shape[0], -1, -1)
    d0, i0 = _ops.rasterize(cu(v), vi_b, H, W, algo=0)
    d1, i1 = _ops.rasterize(cu(v), vi_b, H, W, algo=1)
    assert th.equal(i0, i1) and th.equal(d0.view(th.int32), d1.view(th.int32))


@needs_ref
@pytest.mark.parametrize("scene", RASTER, ids=RASTER_IDS)
def test_rasterize_bit_exact_vs_reference_cuda(scene):
    _, v, vi, H, W = scene
    d_ref, i_ref = R.rasterize_with_depth(cu(v), cu(vi), H, W)
    d, i = drtk_b200.rasterize_with_depth(cu(v), cu(vi), H, W)
    assert th.equal(i, i_ref), f"{int((i != i_ref).sum())} index mismatches"
    assert th.equal(d.view(th.int32), d_ref.view(th.int32)), "depth bits differ"


@needs_ref
@pytest.mark.parametrize("cfg,N,overdraw", [(3, 8, False), (3, 2, True), (4, 2, False), (4, 1, True), (5, 1, False)])
def test_rasterize_bit_exact_vs_reference_cuda_baseline_sizes(cfg, N, overdraw):
    v, vi, H, W = scenes.config_mesh(cfg, N=N, overdraw=overdraw, device=DEV)
    d_ref, i_ref = R.rasterize_with_depth(v, vi, H, W)
    d, i = drtk_b200.rasterize_with_depth(v, vi, H, W)
    assert th.equal(i, i_ref), f"{int((i != i_ref).sum())} index mismatches"
    assert th.equal(d.view(th.int32), d_ref.view(th.int32))


def mixed_size_scene():
    """One 1280 x 1536 image holding all three bin classes of the tiler: a fine mesh (small: <= 2x2 tiles), a 70-px mesh
    (medium: super-tile lists) and a 3x3 mesh of ~600-px cells plus two canvas-sized triangles (large: per-image list),
    interleaved in depth so that every class wins some pixels."""
    H, W = 1280, 1536
    parts, faces, off = [], [], 0
    for nx, ny, z0, z1, seed in ((120, 100, 1.8, 3.2, 31), (22, 18, 1.5, 3.5, 32), (3, 3, 1.2, 3.8, 33)):
        v, vi = scenes.grid_mesh(nx, ny, H, W, 1, seed=seed)
        g = th.Generator().manual_seed(seed)
        v[0, :, 2] = z0 + (z1 - z0) * th.rand((v.shape[1],), generator=g)
        parts.append(v[0]); faces.append(vi + off); off += v.shape[1]
    big = th.tensor([[-200.0, -100.0, 2.5], [1700.0, 30.0, 2.6], [700.0, 1500.0, 2.4],
                     [40.0, 1200.0, 2.2], [1500.0, 1250.0, 3.0], [800.0, -300.0, 2.0]])
    parts.append(big); faces.append(th.tensor([[0, 1, 2], [3, 4, 5]], dtype=th.int32) + off)
    return th.cat(parts)[None].contiguous(), th.cat(faces).contiguous(), H, W


def test_rasterize_triangle_size_classes_agree_bitwise():
    v, vi, H, W = mixed_size_scene()
    vi_b = cu(vi)[None]
    d0, i0 = _ops.rasterize(cu(v), vi_b, H, W, algo=0)
    d1, i1 = _ops.rasterize(cu(v), vi_b, H, W, algo=1)
    assert th.equal(i0, i1) and th.equal(d0.view(th.int32), d1.view(th.int32))
    nf = [2 * 119 * 99, 2 * 21 * 17, 2 * 2 * 2, 2]  # faces per class, in list order
    lo = 0
    for k, n in enumerate(nf):  # every class owns pixels in the result
        assert int(((i0 >= lo) & (i0 < lo + n)).sum()) > 1000, f"class {k} invisible"
        lo += n


@needs_ref
def test_rasterize_triangle_size_classes_vs_reference_cuda():
    v, vi, H, W = mixed_size_scene()
    d_ref, i_ref = R.rasterize_with_depth(cu(v), cu(vi), H, W)
    d, i = drtk_b200.rasterize_with_depth(cu(v), cu(vi), H, W)
    assert th.equal(i, i_ref), f"{int((i != i_ref).sum())} index mismatches"
    assert th.equal(d.view(th.int32), d_ref.view(th.int32)), "depth bits differ"


def test_rasterize_known_answers():
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        ka = json.load(f)
    for key, (v, vi, H, W) in (("hello_triangle_512", scenes.hello_triangle(DEV)), ("two_triangles_512", scenes.two_triangles(DEV))):
        index = npy(drtk_b200.rasterize(v, vi, H, W))
        assert int((index >= 0).sum()) == ka[key]["covered"]
        assert zlib.crc32(index.tobytes()) == ka[key]["index_crc32"]  # integer coordinates: exact in any arithmetic


def test_rasterize_properties_full_size():
    """Config 4 geometry (100k triangles, 2048^2): size-independent properties."""
    v, vi, H, W = scenes.config_mesh(4, N=2, device=DEV)
    depth, index = drtk_b200.rasterize_with_depth(v, vi, H, W)
    F = vi.shape[0]
    assert int(index.min()) >= -1 and int(index.max()) < F
    assert bool(((index == -1) == (depth == 0)).all())
    # idempotence / determinism
    d2, i2 = drtk_b200.rasterize_with_depth(v, vi, H, W)
    assert th.equal(index, i2) and th.equal(depth, d2)
    # watertight single sheet: the mesh interior has no holes -> every pixel strictly inside the
    # hull of the jittered grid is covered; check the central 80 % of the canvas
    y0, y1, x0, x1 = int(0.1 * H), int(0.9 * H), int(0.1 * W), int(0.9 * W)
    assert bool((index[:, y0:y1, x0:x1] >= 0).all())
    # each covered pixel lies inside (or on the boundary of) its triangle: render's barycentrics
    _, bary = drtk_b200.render(v, vi, index)
    assert float(bary.min()) > -1e-4
    s = bary.sum(1)[index >= 0]
    assert float((s - 1).abs().max()) < 1e-5
    # a permutation of the triangle list relabels the ids but not coverage or depth
    perm = th.randperm(F, device=DEV, generator=th.Generator(device=DEV).manual_seed(1))
    d3, i3 = drtk_b200.rasterize_with_depth(v, vi[perm], H, W)
    assert th.equal(d3, depth)
    assert th.equal(th.where(i3 >= 0, perm[i3.clamp(min=0).long()].int(), i3), index)


def test_rasterize_inputs_layouts_and_edge_cases():
    v, vi = scenes.grid_mesh(9, 9, 64, 64, 2, seed=3)
    ref_d, ref_i = _ops.rasterize(cu(v), cu(vi)[None].expand(2, -1, -1), 64, 64)
    # materialised [N,F,3] topology, non-contiguous vertex tensor, top nibble ignored (reference :74)
    vi_b = cu(vi)[None].repeat(2, 1, 1)
    big = th.zeros((2, v.shape[1], 7), device=DEV)
    big[..., 1:6:2] = cu(v)
    v_nc = big[..., 1:6:2]
    assert not v_nc.is_contiguous()
    vi_flag = vi_b.clone()
    vi_flag[..., 0] |= (0x5 << 28)
    for a, b in ((cu(v), vi_b), (v_nc, vi_b), (cu(v), vi_flag)):
        d, i = _ops.rasterize(a, b, 64, 64)
        assert th.equal(i, ref_i) and th.equal(d, ref_d)
    # empty topology and empty batch
    d, i = _ops.rasterize(cu(v), th.zeros((2, 0, 3), dtype=th.int32, device=DEV), 16, 16)
    assert bool((i == -1).all()) and bool((d == 0).all())
    d, i = _ops.rasterize(th.zeros((0, 4, 3), device=DEV), th.zeros((0, 2, 3), dtype=th.int32, device=DEV), 16, 16)
    assert i.shape == (0, 16, 16)
    # degenerate / behind-camera / off-screen triangles draw nothing
    vv = cu(np.array([[[1, 1, 1], [5, 1, 1], [1, 5, 1], [1, 1, -1], [-9, -9, 1], [-5, -9, 1], [-9, -5, 1]]], np.float32))
    for tri in ([0, 0, 0], [0, 1, 3], [4, 5, 6], [0, 1, 1]):
        d, i = _ops.rasterize(vv, cu(np.array([[tri]], np.int32)), 8, 8)
        assert bool((i == -1).all())
    with pytest.raises(RuntimeError, match="int32"):
        drtk_b200.rasterize(cu(v), cu(vi).long(), 8, 8)


# ------------------------------------------------------------------------------------------------
# wireframe rasterisation (src/rasterize/rasterize_kernel.cu:171-400): the reference has no CPU twin for it
# (rasterize_kernel_cpu.cpp:257 raises), so parity is pinned on the reference CUDA kernel: live when
# oracle/_ref travelled to the box, and through tests/golden/wire_*.npz (written by that kernel on a B200,
# tests/golden/make_golden_wireframe.py).
# ------------------------------------------------------------------------------------------------
def with_edge_flags(vi, seed):
    """Edge-visibility nibbles in bits 28-30 of vi[..., 0] ONLY (:293-303): random, or all visible (seed None)."""
    g = th.Generator().manual_seed(seed or 0)
    flags = th.randint(0, 8, (vi.shape[0],), generator=g, dtype=th.int64)
    if seed is None:
        flags = th.full_like(flags, 7)
    out = vi.clone().to(th.int64)
    out[:, 0] = out[:, 0] | (flags << 28)
    return out.to(th.int32)


@needs_ref
@pytest.mark.parametrize("scene", SMALL, ids=SMALL_IDS)
@pytest.mark.parametrize("all_edges", [True, False])
def test_wireframe_bit_exact_vs_reference_cuda(scene, all_edges):
    _, v, vi, H, W = scene
    vi = vi.clone()
    vif = with_edge_flags(vi, None if all_edges else 5)
    d, i = drtk_b200.rasterize_with_depth(cu(v), cu(vif), H, W, wireframe=True)
    dr, ir = R.rasterize_with_depth(cu(v), cu(vif), H, W, wireframe=True)
    assert int((i != ir).sum()) == 0
    assert int((d.view(th.int32) != dr.view(th.int32)).sum()) == 0
    assert int((i >= 0).sum()) > 0  # lines were drawn


@needs_ref
@pytest.mark.parametrize("config,N", [(3, 2), (4, 1)])
def test_wireframe_bit_exact_vs_reference_cuda_baseline_sizes(config, N):
    v, vi, H, W = scenes.config_mesh(config, N=N)
    vif = with_edge_flags(vi, 9)
    d, i = drtk_b200.rasterize_with_depth(cu(v), cu(vif), H, W, wireframe=True)
    dr, ir = R.rasterize_with_depth(cu(v), cu(vif), H, W, wireframe=True)
    assert int((i != ir).sum()) == 0 and int((d.view(th.int32) != dr.view(th.int32)).sum()) == 0


def test_wireframe_golden():
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, "wire_*.npz")))
    assert files, "tests/golden/wire_*.npz missing"
    for fn in files:
        z = np.load(fn)
        d, i = drtk_b200.rasterize_with_depth(cu(z["v"]), cu(z["vi"]), int(z["H"]), int(z["W"]), wireframe=True)
        assert (npy(i) == z["index_img"]).all(), fn
        assert (npy(d).view(np.int32) == z["depth_img"].view(np.int32)).all(), fn


def test_wireframe_properties():
    # no visible edge -> nothing is drawn but the interior still occludes (index -1 everywhere, depth written)
    v, vi, H, W = scenes.two_triangles()
    d, i = drtk_b200.rasterize_with_depth(cu(v), cu(vi), H, W, wireframe=True)
    assert bool((i == -1).all()) and int((d > 0).sum()) > 1000
    # all edges visible: line pixels carry the triangle id; border pixels (x or y = 0 / max) are never touched (:333-337)
    vif = with_edge_flags(vi, None)
    d, i = drtk_b200.rasterize_with_depth(cu(v), cu(vif), H, W, wireframe=True)
    n_line = int((i >= 0).sum())
    assert 500 < n_line < 20000
    assert bool((i[:, 0, :] == -1).all() and (i[:, -1, :] == -1).all() and (i[:, :, 0] == -1).all() and (i[:, :, -1] == -1).all())
    # deterministic
    d2, i2 = drtk_b200.rasterize_with_depth(cu(v), cu(vif), H, W, wireframe=True)
    assert bool((i == i2).all()) and bool((d.view(th.int32) == d2.view(th.int32)).all())


# ------------------------------------------------------------------------------------------------
# render
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_render_golden(name):
    g = load_golden(name)
    v = cu(g["v"]).requires_grad_(True)
    depth, bary = drtk_b200.render(v, cu(g["vi"]), cu(g["index_img"]))
    assert_close(npy(depth), g["depth_img"], what="depth_img")
    assert_close(npy(bary), g["bary_img"], what="bary_img")
    ((bary * cu(g["w_bary"])).sum() + (depth * cu(g["w_depth"])).sum()).backward()
    assert_close(npy(v.grad), g["grad_v_render"], rtol=2e-5, what="grad_v (render)")


@pytest.mark.parametrize("scene", SMALL, ids=SMALL_IDS)
def test_render_vs_oracle(scene):
    _, v, vi, H, W = scene
    _, index = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=1)
    d_o, b_o = O.render_fwd(v.numpy(), vi.numpy(), index)
    vv = cu(v).requires_grad_(True)
    depth, bary = drtk_b200.render(vv, cu(vi), cu(index))
    assert_close(npy(depth), d_o, what="depth_img")
    assert_close(npy(bary), b_o, what="bary_img")
    gen = th.Generator().manual_seed(3)
    wb, wd = th.rand(bary.shape, generator=gen), th.rand(depth.shape, generator=gen)
    ((bary * cu(wb)).sum() + (depth * cu(wd)).sum()).backward()
    g64 = O.render_bwd(v.double().numpy(), vi.numpy(), index, wd.double().numpy(), wb.double().numpy())
    assert_close(npy(vv.grad), g64, rtol=2e-5, what="grad_v vs fp64 oracle")
    # only one of the two upstream gradients defined
    vv.grad = None
    depth2, bary2 = drtk_b200.render(vv, cu(vi), cu(index))
    (depth2 * cu(wd)).sum().backward()
    g64d = O.render_bwd(v.double().numpy(), vi.numpy(), index, wd.double().numpy(), None)
    assert_close(npy(vv.grad), g64d, rtol=2e-5, what="grad_v (depth only)")


def test_render_strided_index_and_no_grad_path():
    v, vi = scenes.grid_mesh(9, 9, 40, 44, 2, seed=5)
    _, index = O.rasterize(v.numpy(), vi.numpy(), 40, 44, mode=1)
    d_o, b_o = O.render_fwd(v.numpy(), vi.numpy(), index)
    wide = th.full((2, 40, 50), -1, dtype=th.int32, device=DEV)
    wide[:, :, 3:47] = cu(index)
    idx_nc = wide[:, :, 3:47]  # row pitch 50, offset 3 -> the scalar path
    depth, bary = drtk_b200.render(cu(v), cu(vi), idx_nc)
    assert_close(npy(bary), b_o, what="bary (strided index)")
    assert not bary.requires_grad and not depth.requires_grad


# ------------------------------------------------------------------------------------------------
# interpolate
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_interpolate_golden(name):
    g = load_golden(name)
    attr = cu(g["attr"]).requires_grad_(True)
    bary = cu(g["bary_img"]).requires_grad_(True)
    out = drtk_b200.interpolate(attr, cu(g["vi"]), cu(g["index_img"]), bary)
    assert_close(npy(out), g["interp"], what="interpolate")
    (out * cu(g["w_img"])).sum().backward()
    assert_close(npy(attr.grad), g["grad_attr"], rtol=2e-5, what="vert_attributes_grad")
    assert_close(npy(bary.grad), g["grad_bary"], what="bary_img_grad")


# C % 4 == 0 -> quad-walker kernel (4/12: partly idle walker lanes; 20/32/36: several channel passes per tile);
# other C -> lane-per-channel tile kernel
@pytest.mark.parametrize("C", [1, 2, 3, 4, 5, 8, 12, 16, 19, 20, 32, 36])
@pytest.mark.parametrize("H,W", [(96, 128), (70, 260), (61, 97)])  # tiled path (two sizes) / generic path
def test_interpolate_vs_oracle_channels(C, H, W):
    v, vi = scenes.grid_mesh(17, 13, H, W, 2, seed=31, overdraw=True)
    _, index = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=1)
    _, bary = O.render_fwd(v.numpy(), vi.numpy(), index)
    attr = scenes.vertex_attributes(2, v.shape[1], C, seed=C)
    out_o = O.interpolate_fwd(attr.numpy(), vi.numpy(), index, bary)
    a = cu(attr).requires_grad_(True)
    b = cu(bary).requires_grad_(True)
    out = drtk_b200.interpolate(a, cu(vi), cu(index), b)
    assert_close(npy(out), out_o, what=f"interpolate C={C}")
    gen = th.Generator().manual_seed(C)
    w = th.rand(out.shape, generator=gen)
    (out * cu(w)).sum().backward()
    ga64, gb64 = O.interpolate_bwd(w.double().numpy(), attr.double().numpy(), vi.numpy(), index, bary.astype(np.float64))
    assert_close(npy(a.grad), ga64, rtol=2e-5, what=f"attr grad C={C}")
    assert_close(npy(b.grad), gb64, what=f"bary grad C={C}")
    # only one input requires grad (the reference instantiates <bary,vert> = <1,0>, <0,1>, <1,1>)
    a2 = cu(attr).requires_grad_(True)
    out2 = drtk_b200.interpolate(a2, cu(vi), cu(index), cu(bary))
    out2.sum().backward()  # expanded (stride-0) upstream gradient
    ga_ones, _ = O.interpolate_bwd(np.ones(out_o.shape), attr.double().numpy(), vi.numpy(), index, bary.astype(np.float64), True, False)
    assert_close(npy(a2.grad), ga_ones, rtol=2e-5, what=f"attr grad (expanded grad_out) C={C}")
    b3 = cu(bary).requires_grad_(True)
    out3 = drtk_b200.interpolate(cu(attr), cu(vi), cu(index), b3)
    (out3 * cu(w)).sum().backward()
    assert_close(npy(b3.grad), gb64, what=f"bary grad only C={C}")


def test_interpolate_backward_batched_topology_and_strides():
    """Per-image topology (vi [N,F,3] with different faces per image -> per-image triangle table) and a
    channel-sliced (non-contiguous) attribute table / upstream gradient through the tiled backward."""
    H, W, C, N = 64, 96, 8, 3
    v, vi = scenes.grid_mesh(11, 9, H, W, N, seed=77, overdraw=True)
    g = th.Generator().manual_seed(5)
    vib = th.stack([vi[th.randperm(vi.shape[0], generator=g)] for _ in range(N)])  # different face order per image
    _, index = O.rasterize(v.numpy(), vib.numpy(), H, W, mode=1)
    _, bary = O.render_fwd(v.numpy(), vib.numpy(), index)
    attr = scenes.vertex_attributes(N, v.shape[1], C, seed=3)
    w = th.rand((N, C, H, W), generator=g)
    ga64, gb64 = O.interpolate_bwd(w.double().numpy(), attr.double().numpy(), vib.numpy(), index, bary.astype(np.float64))
    a = cu(attr).requires_grad_(True)
    b = cu(bary).requires_grad_(True)
    out = drtk_b200.interpolate(a, cu(vib), cu(index), b)
    (out * cu(w)).sum().backward()
    assert_close(npy(a.grad), ga64, rtol=2e-5, what="attr grad (batched vi)")
    assert_close(npy(b.grad), gb64, what="bary grad (batched vi)")
    # attribute table that is a channel slice of a wider tensor (row stride 2C): falls off the 128-bit row path
    wide = th.zeros((N, v.shape[1], 2 * C), device=DEV)
    wide[..., :C] = cu(attr)
    a2 = wide[..., :C].detach().requires_grad_(True)
    b2 = cu(bary).requires_grad_(True)
    out2 = drtk_b200.interpolate(a2, cu(vib), cu(index), b2)
    (out2 * cu(w)).sum().backward()
    assert_close(npy(a2.grad), ga64, rtol=2e-5, what="attr grad (strided attributes)")
    assert_close(npy(b2.grad), gb64, what="bary grad (strided attributes)")


def test_interpolate_background_sweep_and_layouts():
    v, vi = scenes.grid_mesh(5, 5, 30, 37, 1, seed=41)  # odd width -> scalar pixel path
    _, index = O.rasterize(v.numpy(), vi.numpy(), 30, 37, mode=1)
    _, bary = O.render_fwd(v.numpy(), vi.numpy(), index)
    attr = scenes.vertex_attributes(1, 25, 6, seed=2)
    out_o = O.interpolate_fwd(attr.numpy(), vi.numpy(), index, bary)
    out = drtk_b200.interpolate(cu(attr), cu(vi), cu(index), cu(bary))
    assert (index == -1).any()
    assert_close(npy(out), out_o, what="interpolate (odd width, background sweep)")
    # channel-sliced (non-contiguous) attribute table
    wide = th.zeros((1, 25, 12), device=DEV)
    wide[..., ::2] = cu(attr)
    out2 = drtk_b200.interpolate(wide[..., ::2], cu(vi), cu(index), cu(bary))
    assert_close(npy(out2), out_o, what="interpolate (strided attributes)")
    with pytest.raises(RuntimeError, match="same dtype"):
        drtk_b200.interpolate(cu(attr), cu(vi), cu(index), cu(bary).double())


# ------------------------------------------------------------------------------------------------
# sparse interpolation matrices (src/interpolate/interpolate_kernel.cu:301-452, interpolate_module.cpp:28-308)
# ------------------------------------------------------------------------------------------------
def _mat_files():
    import glob
    return sorted(glob.glob(os.path.join(GOLDEN, "mat_*.npz")))


def test_interpolation_matrix_golden():
    files = _mat_files()
    assert len(files) >= 2
    for fn in files:
        z = np.load(fn)
        V = int(z["V"])
        b = cu(z["bary_img"]).requires_grad_(True)
        A = drtk_b200.interpolation_matrix(cu(z["vi"]), cu(z["index_img"]), b, V)
        assert A.layout == th.sparse_csr and tuple(A.shape) == (z["row_pixels"].size, V)
        assert (npy(A.crow_indices()) == z["crow"]).all() and (npy(A.col_indices()) == z["col"]).all()
        assert (npy(A.values()) == z["values"]).all()
        (A.values() * cu(z["w_values"])).sum().backward()
        assert (npy(b.grad) == z["grad_bary"]).all()
        # pixel_values = A @ X reproduces interpolate() at the foreground pixels
        N, H, W = z["index_img"].shape
        X = th.rand((V, 5), generator=th.Generator().manual_seed(1))
        px = (A.detach() @ cu(X))
        img = drtk_b200.interpolate(cu(X)[None].expand(N, -1, -1).contiguous(), cu(z["vi"]), cu(z["index_img"]), cu(z["bary_img"]))
        ref = img.permute(0, 2, 3, 1).reshape(-1, 5)[cu(z["row_pixels"])]
        assert_close(npy(px), npy(ref), what="A @ X vs interpolate")


def test_interpolation_normal_matrix_golden_and_cache():
    for fn in _mat_files():
        z = np.load(fn)
        V = int(z["V"])
        vi = cu(z["vi"])
        b = cu(z["bary_img"]).requires_grad_(True)
        M = drtk_b200.interpolation_normal_matrix(vi, cu(z["index_img"]), b, V)
        assert M.layout == th.sparse_csr and tuple(M.shape) == (V, V)
        assert (npy(M.crow_indices()) == z["n_crow"]).all() and (npy(M.col_indices()) == z["n_col"]).all()
        assert_close(npy(M.values()), z["n_values"], rtol=2e-5, what="normal matrix values")
        (M.values() * cu(z["n_w_values"])).sum().backward()
        assert_close(npy(b.grad), z["n_grad_bary"], rtol=2e-5, what="normal matrix bary grad")
        # the topology structure is cached per vi tensor (identity + version): second call reuses the same buffers
        M2 = drtk_b200.interpolation_normal_matrix(vi, cu(z["index_img"]), cu(z["bary_img"]), V)
        assert M2.crow_indices().data_ptr() == M.crow_indices().data_ptr()
        vi.add_(0)  # in-place edit bumps the version counter -> rebuild
        M3 = drtk_b200.interpolation_normal_matrix(vi, cu(z["index_img"]), cu(z["bary_img"]), V)
        assert M3.crow_indices().data_ptr() != M.crow_indices().data_ptr()
        assert (npy(M3.col_indices()) == z["n_col"]).all()
    with pytest.raises(RuntimeError, match="outside"):
        drtk_b200.interpolation_normal_matrix(cu(np.array([[0, 1, 9]], np.int32)), cu(np.full((1, 4, 4), -1, np.int32)),
                                              th.zeros((1, 3, 4, 4), device=DEV), 4)


@needs_ref
@pytest.mark.parametrize("scene", SMALL[:3], ids=SMALL_IDS[:3])
def test_interpolation_matrices_vs_reference_cuda(scene):
    R.load()
    _, v, vi, H, W = scene
    N, V = v.shape[0], v.shape[1]
    vin = cu(vi)[None].expand(N, -1, -1).contiguous()
    index = drtk_b200.rasterize(cu(v), vin, H, W)
    _, bary = drtk_b200.render(cu(v), vin, index)
    g = th.Generator(device=DEV).manual_seed(3)
    # A
    b0, b1 = bary.clone().requires_grad_(True), bary.clone().requires_grad_(True)
    A = drtk_b200.interpolation_matrix(vin, index, b0, V)
    crow, col, val, rows = th.ops.interpolate_ext.interpolation_matrix(vin, index, b1)
    assert bool((A.crow_indices() == crow).all()) and bool((A.col_indices() == col).all()) and bool((A.values() == val).all())
    w = th.rand(val.shape, device=DEV, generator=g)
    (A.values() * w).sum().backward(); (val * w).sum().backward()
    assert bool((b0.grad == b1.grad).all())
    # A^T A
    b2, b3 = bary.clone().requires_grad_(True), bary.clone().requires_grad_(True)
    M = drtk_b200.interpolation_normal_matrix(vin, index, b2, V)
    ncrow, ncol, nval = th.ops.interpolate_ext.interpolation_normal_matrix(vin, index, b3, V)
    assert bool((M.crow_indices() == ncrow).all()) and bool((M.col_indices() == ncol).all())
    assert_close(npy(M.values()), npy(nval), rtol=2e-5, what="normal matrix values vs reference CUDA")
    w2 = th.rand(nval.shape, device=DEV, generator=g)
    (M.values() * w2).sum().backward(); (nval * w2).sum().backward()
    assert_close(npy(b2.grad), npy(b3.grad), rtol=2e-5, what="normal matrix bary grad vs reference CUDA")
    # against the dense oracle
    o = O.interpolation_normal_matrix(npy(vin), npy(index), npy(bary), V)
    assert_close(npy(M.to_dense()), o["dense"], rtol=2e-5, what="A^T A dense")


# ------------------------------------------------------------------------------------------------
# edge_grad_estimator
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_edge_grad_backward_golden(name):
    g = load_golden(name)
    out = _ops.edge_grad_backward(cu(g["v"]), cu(g["interp"]), cu(g["index_img"]),
                                  cu(g["vi"])[None].expand(g["v"].shape[0], -1, -1), cu(g["w_img"]), 1e4)
    assert_close(npy(out), g["grad_v_pix_img"], what="grad_v_pix_img")


@pytest.mark.parametrize("scene", SMALL, ids=SMALL_IDS)
@pytest.mark.parametrize("max_dp_dr", [1e4, 0.0])
def test_edge_grad_backward_vs_oracle(scene, max_dp_dr):
    _, v, vi, H, W = scene
    N, V = v.shape[:2]
    _, index = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=1)
    gen = th.Generator().manual_seed(9)
    img, go = th.rand((N, 4, H, W), generator=gen), th.randn((N, 4, H, W), generator=gen)
    ref = O.edge_grad_bwd(v.numpy(), img.numpy(), index, vi.numpy(), go.numpy(), max_dp_dr)
    out = _ops.edge_grad_backward(cu(v), cu(img), cu(index), cu(vi)[None].expand(N, -1, -1), cu(go), max_dp_dr)
    a, e = npy(out), ref
    # the clamp-free intersection branch divides by sin(angle between normals): compare those few
    # ill-conditioned pixels relative to their own magnitude, everything else at the 1e-5 bar
    big = np.abs(e) > 1e3 * np.median(np.abs(e[e != 0])) if (e != 0).any() else np.zeros_like(e, bool)
    assert_close(np.where(big, 0, a), np.where(big, 0, e), what="grad_v_pix_img")
    if big.any():
        assert np.all(np.abs(a[big] - e[big]) <= 1e-3 * np.abs(e[big]))


@needs_ref
@pytest.mark.parametrize("cfg,N,overdraw", [(3, 2, False), (3, 1, True)])
def test_pipeline_vs_reference_cuda(cfg, N, overdraw):
    """Full forward+backward pipeline on a BASELINE-size mesh against the reference's own CUDA kernels."""
    v, vi, H, W = scenes.config_mesh(cfg, N=N, overdraw=overdraw, device=DEV)
    C = 16
    attr = scenes.vertex_attributes(N, v.shape[1], C, seed=77, device=DEV)
    w = th.rand((N, C, H, W), device=DEV, generator=th.Generator(device=DEV).manual_seed(5))
    res = {}
    for tag, api in (("ref", R), ("new", drtk_b200)):
        vv, aa = v.clone().requires_grad_(True), attr.clone().requires_grad_(True)
        cap = {}
        index = api.rasterize(vv, vi, H, W)
        depth, bary = api.render(vv, vi, index)
        img = api.interpolate(aa, vi, index, bary)
        img = api.edge_grad_estimator(vv, vi, bary, img, index, v_pix_img_hook=lambda g: cap.__setitem__("g", g.clone()))
        ((img * w).sum() + depth.sum()).backward()
        res[tag] = dict(index=index, depth=depth.detach(), bary=bary.detach(), img=img.detach(), gv=vv.grad, ga=aa.grad, gpix=cap["g"])
    r, n = res["ref"], res["new"]
    assert th.equal(r["index"], n["index"])
    assert_close(npy(n["depth"]), npy(r["depth"]), what="depth")
    assert_close(npy(n["bary"]), npy(r["bary"]), what="bary")
    assert_close(npy(n["img"]), npy(r["img"]), what="img")
    assert_close(npy(n["gpix"]), npy(r["gpix"]), what="grad_v_pix_img")
    # vertex gradients are long atomically-ordered fp32 sums in the reference: compare both with
    # slack for that ordering noise (each is ~1e-6 of scale away from the exact sum)
    assert_close(npy(n["ga"]), npy(r["ga"]), rtol=5e-5, what="grad attr")
    assert_close(npy(n["gv"]), npy(r["gv"]), rtol=5e-5, what="grad v")


@needs_ref
@pytest.mark.parametrize("overdraw,hook", [(False, False), (False, True), (True, False), (True, True)])
def test_pipeline_config4_batch8_vs_reference_cuda(overdraw, hook):
    """The MEASURED configuration (BASELINE config 4: 100 352 triangles, 2048^2, batch 8, C = 16, bench.py's seeds:
    mesh 4000 + b, attributes 4001, cotangent 4002) and its overdraw-2 variant (occlusion + intersections: the
    horiz_int / vert_int branches of src/edge_grad/edge_grad_kernel.cu:320-341 at size), with and without a
    v_pix_img hook (the hook selects the two-kernel edge_grad plan, no hook the fused one), against the reference's
    CUDA kernels.  Compared on the device: index_img equal, fp32 outputs 1e-5, vertex gradients 5e-5 of scale."""
    cfg, N, C = 4, 8, 16
    c = scenes.CONFIGS[cfg]
    H, W = c["H"], c["W"]
    v, vi = scenes.grid_mesh(c["nx"], c["ny"], H, W, N, seed=1000 * cfg, overdraw=overdraw, device=DEV)
    attr = scenes.vertex_attributes(N, v.shape[1], C, seed=1000 * cfg + 1, device=DEV)
    w = th.rand((N, C, H, W), device=DEV, generator=th.Generator(device=DEV).manual_seed(1000 * cfg + 2))
    res = {}
    for tag, api in (("ref", R), ("new", drtk_b200)):
        vv, aa = v.clone().requires_grad_(True), attr.clone().requires_grad_(True)
        cap = {}
        index = api.rasterize(vv, vi, H, W)
        depth, bary = api.render(vv, vi, index)
        img = api.interpolate(aa, vi, index, bary)
        out = api.edge_grad_estimator(vv, vi, bary, img, index,
                                      v_pix_img_hook=(lambda g: cap.__setitem__("g", g.clone())) if hook else None)
        out.backward(gradient=w)
        res[tag] = dict(index=index, depth=depth.detach(), bary=bary.detach(), img=img.detach(), gv=vv.grad, ga=aa.grad,
                        gpix=cap.get("g"))
        del out, img, bary, depth
    r, n = res["ref"], res["new"]
    assert th.equal(r["index"], n["index"]), f"index_img: {int((r['index'] != n['index']).sum())} pixels differ"
    if overdraw:  # the scene really has occlusion: both sheets visible somewhere
        F1 = vi.shape[0] // 2
        assert bool((n["index"] >= F1).any()) and bool(((n["index"] >= 0) & (n["index"] < F1)).any())
    assert_close_device(n["depth"], r["depth"], 1e-5, "depth_img")
    assert_close_device(n["bary"], r["bary"], 1e-5, "bary_img")
    assert_close_device(n["img"], r["img"], 1e-5, "interpolated img")
    if hook:
        assert_edge_gradient_image_close(n["gpix"], r["gpix"], "grad_v_pix_img")
    assert_close_device(n["ga"], r["ga"], 5e-5, "grad attr")
    assert_close_device(n["gv"], r["gv"], 5e-5, "grad v")


def test_empty_face_list():
    """F == 0 (the reference handles it: background sweep forward, zero gradients backward)."""
    N, V, C, H, W = 2, 5, 4, 16, 24
    g = th.Generator().manual_seed(1)
    v = cu(th.rand((N, V, 3), generator=g) * 10 + 1)
    vi = th.zeros((0, 3), dtype=th.int32, device=DEV)
    attr = cu(th.rand((N, V, C), generator=g)).requires_grad_(True)
    index = drtk_b200.rasterize(v, vi, H, W)
    assert bool((index == -1).all())
    _, bary = drtk_b200.render(v, vi, index)
    bary = bary.requires_grad_(True)
    img = drtk_b200.interpolate(attr, vi, index, bary)
    xs = (2 * th.arange(W, device=DEV) + 1) / W - 1
    ys = (2 * th.arange(H, device=DEV) + 1) / H - 1
    assert th.allclose(img[:, 0], xs[None, None, :].expand(N, H, W)) and th.allclose(img[:, 1], ys[None, :, None].expand(N, H, W))
    img.sum().backward()
    assert bool((attr.grad == 0).all()) and bool((bary.grad == 0).all())  # not uninitialised memory
    ga, gb = _ops.interpolate_backward(th.ones_like(img), attr.detach(), vi[None].expand(N, -1, -1), index, bary.detach(), True, True)
    assert bool((ga == 0).all()) and bool((gb == 0).all())


@pytest.mark.parametrize("name", CASES)
def test_pipeline_autograd_golden(name):
    """The drop-in API end to end: same autograd surface as the reference (hook, identity forward,
    output requires grad), gradients equal to the reference's."""
    g = load_golden(name)
    H, W = map(int, g["HW"])
    v = cu(g["v"]).requires_grad_(True)
    attr = cu(g["attr"]).requires_grad_(True)
    vi = cu(g["vi"])
    index = drtk_b200.rasterize(v, vi, H, W)
    np.testing.assert_array_equal(npy(index), g["index_img"])
    assert not index.requires_grad
    depth, bary = drtk_b200.render(v, vi, index)
    img = drtk_b200.interpolate(attr, vi, index, bary)
    seen = {}
    out = drtk_b200.edge_grad_estimator(v, vi, bary, img, index, v_pix_img_hook=lambda gr: seen.__setitem__("g", gr.clone()))
    assert out.requires_grad and th.equal(out, img)
    (out * cu(g["w_img"])).sum().backward()
    assert_close(npy(seen["g"]), g["grad_v_pix_img"], what="hooked grad_v_pix_img")
    assert_close(npy(v.grad), g["grad_v_full"], rtol=2e-5, what="grad_v (pipeline)")
    assert_close(npy(attr.grad), g["grad_attr_full"], rtol=2e-5, what="grad_attr (pipeline)")


@pytest.mark.parametrize("name", CASES)
def test_pipeline_autograd_golden_no_hook(name):
    """Without a hook the fused edge_grad+conduit kernel runs; the vertex gradients are the same."""
    g = load_golden(name)
    H, W = map(int, g["HW"])
    v = cu(g["v"]).requires_grad_(True)
    attr = cu(g["attr"]).requires_grad_(True)
    vi = cu(g["vi"])
    index = drtk_b200.rasterize(v, vi, H, W)
    depth, bary = drtk_b200.render(v, vi, index)
    img = drtk_b200.interpolate(attr, vi, index, bary)
    out = drtk_b200.edge_grad_estimator(v, vi, bary, img, index)
    assert out.requires_grad and th.equal(out, img)
    (out * cu(g["w_img"])).sum().backward()
    assert_close(npy(v.grad), g["grad_v_full"], rtol=2e-5, what="grad_v (fused pipeline)")
    assert_close(npy(attr.grad), g["grad_attr_full"], rtol=2e-5, what="grad_attr (fused pipeline)")


@needs_ref
def test_config5_slice_vs_reference_cuda():
    """One image of BASELINE config 5 (1 002 528 triangles, 4096 x 4096, 16 attributes): the whole forward + backward
    against the reference's CUDA kernels, compared on the device (2^30 output elements: 64-bit image offsets)."""
    N, C = 1, 16
    v, vi, H, W = scenes.config_mesh(5, N=N, device=DEV)
    attr = scenes.vertex_attributes(N, v.shape[1], C, seed=5, device=DEV)
    w = th.rand((N, C, H, W), device=DEV, generator=th.Generator(device=DEV).manual_seed(6))
    res = {}
    for tag, api in (("ref", R), ("new", drtk_b200)):
        vv, aa = v.clone().requires_grad_(True), attr.clone().requires_grad_(True)
        index = api.rasterize(vv, vi, H, W)
        depth, bary = api.render(vv, vi, index)
        img = api.interpolate(aa, vi, index, bary)
        out = api.edge_grad_estimator(vv, vi, bary, img, index)
        out.backward(gradient=w)
        res[tag] = (index, depth.detach(), bary.detach(), img.detach(), vv.grad, aa.grad)
        del out, img, bary, depth
    r, n = res["ref"], res["new"]
    assert th.equal(r[0], n[0]), "index_img"
    for k, name, tol in ((1, "depth", 1e-5), (2, "bary", 1e-5), (3, "img", 1e-5), (4, "grad_v", 5e-5), (5, "grad_attr", 5e-5)):
        assert_close_device(n[k], r[k], tol, name)


@pytest.mark.parametrize("cfg,N,overdraw,expand_vi", [(3, 2, False, True), (3, 1, True, False), (4, 1, False, True)])
def test_edge_grad_fused_equals_two_kernel_plan(cfg, N, overdraw, expand_vi):
    """C-ABI level: drtk_b200_edge_grad_backward_fused == edge_grad_backward -> interpolate_backward(C=3)."""
    v, vi, H, W = scenes.config_mesh(cfg, N=N, overdraw=overdraw, device=DEV)
    vi = vi[None].expand(N, -1, -1)  # the _ops launchers take the batched [N,F,3] form
    if not expand_vi:
        vi = vi.contiguous()
    index = drtk_b200.rasterize(v, vi, H, W)
    _, bary = drtk_b200.render(v, vi, index)
    attr = scenes.vertex_attributes(N, v.shape[1], 5, seed=3, device=DEV)
    img = drtk_b200.interpolate(attr, vi, index, bary)
    go = th.randn(img.shape, device=DEV, generator=th.Generator(device=DEV).manual_seed(11))
    gimg = _ops.edge_grad_backward(v, img, index, vi, go, 1e4)
    gv2, _ = _ops.interpolate_backward(gimg, v, vi, index, bary, True, False)
    gv1 = _ops.edge_grad_backward_fused(v, img, index, vi, go, bary, 1e4)
    assert float(gv2.abs().max()) > 0
    assert_close(npy(gv1), npy(gv2), rtol=2e-5, what="fused grad_v_pix")
    # non-contiguous bary / grad_output strides go through the same kernel
    bary_nc = bary.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    go_nc = go.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    gv3 = _ops.edge_grad_backward_fused(v, img, index, vi, go_nc, bary_nc, 1e4)
    assert_close(npy(gv3), npy(gv2), rtol=2e-5, what="fused grad_v_pix (strided)")


def test_edge_grad_estimator_autograd_corner_cases():
    g = load_golden("grid_48")
    H, W = map(int, g["HW"])
    vi, index, bary = cu(g["vi"]), cu(g["index_img"]), cu(g["bary_img"])
    # img without grad history still yields gradients w.r.t. v_pix (reference docstring :45-47)
    v = cu(g["v"]).requires_grad_(True)
    img = cu(g["interp"])
    out = drtk_b200.edge_grad_estimator(v, vi, bary, img, index)
    assert out.requires_grad
    (out * cu(g["w_img"])).sum().backward()
    gv, _ = O.interpolate_bwd(g["grad_v_pix_img"].astype(np.float64), g["v"].astype(np.float64), g["vi"], g["index_img"],
                              g["bary_img"].astype(np.float64), True, False)
    assert_close(npy(v.grad), gv, rtol=2e-5, what="grad_v through edge_grad only")
    # v_pix without grad: the op is a pure pass-through for img's gradient
    img2 = cu(g["interp"]).requires_grad_(True)
    out2 = drtk_b200.edge_grad_estimator(cu(g["v"]), vi, bary, img2, index)
    (out2 * 2).sum().backward()
    assert bool((img2.grad == 2).all())
    # a hook may replace the gradient image (torch hook semantics)
    v3 = cu(g["v"]).requires_grad_(True)
    out3 = drtk_b200.edge_grad_estimator(v3, vi, bary, cu(g["interp"]), index, v_pix_img_hook=lambda gr: gr * 0)
    (out3 * cu(g["w_img"])).sum().backward()
    assert bool((v3.grad == 0).all())


def test_drop_in_fitting_loop_two_triangles():
    """The reference's only test artefact (test/two_triangles.py) as a short, asserted fit."""
    drtk_b200.install_as_drtk()
    import drtk
    v_gt, vi, H, W = scenes.two_triangles(DEV)
    vt = th.zeros(1, 6, 2, device=DEV)
    vt[:, 3:6, 0] = 1

    def shade(v):
        index = drtk.rasterize(v, vi, H, W)
        _, bary = drtk.render(v, vi, index)
        uv = drtk.interpolate(vt, vi, index, bary)
        img = (0.5 + 0.5 * uv[:, :1]) * (index != -1)[:, None]
        return drtk.edge_grad_estimator(v_pix=v, vi=vi, bary_img=bary, img=img, index_img=index)

    with th.no_grad():
        target = shade(v_gt)
    th.manual_seed(10)
    v = th.nn.Parameter(v_gt + th.randn_like(v_gt) * 8.0)
    opt = th.optim.Adam([v], lr=0.5)
    losses = []
    for _ in range(60):
        loss = ((shade(v) - target) ** 2).mean()
        opt.zero_grad()
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.5 * losses[0], losses[::10]


def test_pipeline_under_cuda_graph_capture():
    """The whole step (forward + backward) is capturable in a CUDA graph: no kernel of the path synchronises the
    device or allocates outside the stream-ordered allocators.  Replays with new vertex positions must reproduce the
    eager results bit for bit (vertex gradients: same REDs in a different order -> tolerance)."""
    H, W, N, C = 128, 160, 2, 8
    v0, vi = scenes.grid_mesh(17, 15, H, W, N, seed=41, overdraw=True)
    v1, _ = scenes.grid_mesh(17, 15, H, W, N, seed=42, overdraw=True)
    attr = scenes.vertex_attributes(N, v0.shape[1], C, seed=43, device=DEV)
    w = th.rand((N, C, H, W), device=DEV, generator=th.Generator(device=DEV).manual_seed(44))
    vid = vi.to(DEV)

    def step(v, a):
        index = drtk_b200.rasterize(v, vid, H, W)
        depth, bary = drtk_b200.render(v, vid, index)
        img = drtk_b200.interpolate(a, vid, index, bary)
        img = drtk_b200.edge_grad_estimator(v, vid, bary, img, index)
        gv, ga = th.autograd.grad(img, (v, a), grad_outputs=w)
        return index, img.detach(), gv, ga

    sv = v0.to(DEV).clone().requires_grad_(True)
    sa = attr.clone().requires_grad_(True)
    side = th.cuda.Stream()
    side.wait_stream(th.cuda.current_stream())
    with th.cuda.stream(side):  # warm-up off the capture stream, as torch's graph recipe asks
        for _ in range(2):
            step(sv, sa)
    th.cuda.current_stream().wait_stream(side)
    graph = th.cuda.CUDAGraph()
    with th.cuda.graph(graph):
        outs = step(sv, sa)
    for vsrc in (v1, v0):
        with th.no_grad():
            sv.copy_(vsrc.to(DEV))
        graph.replay()
        th.cuda.synchronize()
        ev = vsrc.to(DEV).clone().requires_grad_(True)
        ea = attr.clone().requires_grad_(True)
        eager = step(ev, ea)
        assert th.equal(outs[0], eager[0]) and th.equal(outs[1], eager[1])
        assert_close(npy(outs[2]), npy(eager[2]), rtol=5e-5, what="grad_v under graph replay")
        assert_close(npy(outs[3]), npy(eager[3]), rtol=5e-5, what="grad_attr under graph replay")


@pytest.mark.parametrize("half", [th.float16, th.bfloat16])
def test_autocast_casts_to_float32_like_the_reference(half):
    """Under torch.autocast the reference's Autocast kernels cast floating inputs to float32 (src/*/..._module.cpp);
    outputs are float32, gradients come back in the leaves' own dtype; outside autocast half inputs are refused."""
    H, W, N, C = 64, 96, 2, 4
    v, vi = scenes.grid_mesh(9, 8, H, W, N, seed=51)
    vh = (v.to(DEV) * th.tensor([1.0, 1.0, 1.0], device=DEV)).to(half)
    ah = scenes.vertex_attributes(N, v.shape[1], C, seed=52, device=DEV).to(half)
    vid = vi.to(DEV)

    def run(vv, aa):
        index = drtk_b200.rasterize(vv, vid, H, W)
        depth, bary = drtk_b200.render(vv, vid, index)
        img = drtk_b200.interpolate(aa, vid, index, bary)
        out = drtk_b200.edge_grad_estimator(vv, vid, bary, img, index)
        (out.sum() + depth.sum()).backward()
        return index, depth, bary, out

    v1, a1 = vh.clone().requires_grad_(True), ah.clone().requires_grad_(True)
    with th.autocast("cuda", dtype=half):
        r1 = run(v1, a1)
    v2, a2 = vh.float().requires_grad_(True), ah.float().requires_grad_(True)
    r2 = run(v2, a2)
    assert r1[1].dtype == r1[2].dtype == r1[3].dtype == th.float32
    for x, y in zip(r1, r2):
        assert th.equal(x, y)
    assert v1.grad.dtype == half and a1.grad.dtype == half
    assert_close(npy(v1.grad.float()), npy(v2.grad), rtol=2e-2, what="grad_v through the half cast")
    assert_close(npy(a1.grad.float()), npy(a2.grad), rtol=2e-2, what="grad_attr through the half cast")
    with pytest.raises(RuntimeError, match="float32 only|same dtype"):
        drtk_b200.render(vh, vid, r2[0])
