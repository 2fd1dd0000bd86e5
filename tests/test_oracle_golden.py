"""CPU suite: the oracle (oracle/drtk_oracle_impl.h) against vectors produced by the reference.

The golden files were written by tests/golden/make_golden.py, which runs the UNMODIFIED reference
kernels (CPU twins compiled from /root/reference by oracle/build_ref.py).  This is what pins the
oracle; the GPU suite then compares the CUDA kernels with the oracle, with the same golden vectors
and -- when oracle/_ref is present on the box -- with the reference's own CUDA kernels.
"""
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from tests.util import GOLDEN, assert_close, golden_cases, load_golden, ulp_diff

CASES = golden_cases()


def test_golden_present():
    assert set(CASES) >= {"two_tri_128", "grid_48", "overdraw_64x48", "fan_32"}


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", [0, 1])
def test_rasterize_matches_reference(name, mode):
    g = load_golden(name)
    H, W = map(int, g["HW"])
    depth, index = O.rasterize(g["v"], g["vi"], H, W, mode=mode)
    # bit-exact triangle ids (integer work), in both the CPU-twin and the CUDA arithmetic
    np.testing.assert_array_equal(index, g["index_img"])
    assert (depth[index < 0] == 0).all()
    assert ulp_diff(depth, g["raster_depth"]).max() <= 8


@pytest.mark.parametrize("name", CASES)
def test_rasterize_f64_agrees(name):
    g = load_golden(name)
    H, W = map(int, g["HW"])
    _, index = O.rasterize(g["v"].astype(np.float64), g["vi"], H, W, mode=0)
    assert (index != g["index_img"]).mean() < 1e-3  # only rounding-level edge pixels may differ


@pytest.mark.parametrize("name", CASES)
def test_render_forward(name):
    g = load_golden(name)
    depth, bary = O.render_fwd(g["v"], g["vi"], g["index_img"])
    assert_close(depth, g["depth_img"], what="depth_img")
    assert_close(bary, g["bary_img"], what="bary_img")


@pytest.mark.parametrize("name", CASES)
def test_render_backward(name):
    g = load_golden(name)
    gv = O.render_bwd(g["v"], g["vi"], g["index_img"], g["w_depth"], g["w_bary"])
    assert_close(gv, g["grad_v_render"], what="grad_v (render)")


@pytest.mark.parametrize("name", CASES)
def test_interpolate_forward_backward(name):
    g = load_golden(name)
    out = O.interpolate_fwd(g["attr"], g["vi"], g["index_img"], g["bary_img"])
    assert_close(out, g["interp"], what="interpolate")
    ga, gb = O.interpolate_bwd(g["w_img"], g["attr"], g["vi"], g["index_img"], g["bary_img"])
    assert_close(ga, g["grad_attr"], what="vert_attributes_grad")
    assert_close(gb, g["grad_bary"], what="bary_img_grad")


@pytest.mark.parametrize("name", CASES)
def test_edge_grad_backward(name):
    g = load_golden(name)
    out = O.edge_grad_bwd(g["v"], g["interp"], g["index_img"], g["vi"], g["w_img"], 1e4)
    assert_close(out, g["grad_v_pix_img"], what="grad_v_pix_img")
    # structural property (reference :270): the last row/column only receive neighbour terms
    # -> no horizontal pair ends in the last row, no vertical pair in the last column
    assert (out[:, 0, -1, :] == 0).all() and (out[:, 1, :, -1] == 0).all()
    assert (g["grad_v_pix_img"][:, 0, -1, :] == 0).all() and (g["grad_v_pix_img"][:, 1, :, -1] == 0).all()


@pytest.mark.parametrize("name", CASES)
def test_full_pipeline_gradient_composes(name):
    """grad_v of the whole pipeline = interpolate-backward(C=3) of grad_v_pix_img + render-backward of
    the bary gradient (the autograd graph of drtk/edge_grad_estimator.py:165-180)."""
    g = load_golden(name)
    gv_edge, _ = O.interpolate_bwd(g["grad_v_pix_img"], g["v"], g["vi"], g["index_img"], g["bary_img"],
                                   need_attr=True, need_bary=False)
    gv_render = O.render_bwd(g["v"], g["vi"], g["index_img"], None, g["grad_bary"])
    assert_close(gv_edge + gv_render, g["grad_v_full"], rtol=2e-5, what="grad_v (pipeline)")
    assert_close(g["grad_attr"], g["grad_attr_full"], what="grad_attr (pipeline)")


def test_known_answers():
    from drtk_b200 import scenes
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        ka = json.load(f)
    for key, scene in (("hello_triangle_512", scenes.hello_triangle()), ("two_triangles_512", scenes.two_triangles())):
        v, vi, H, W = scene
        depth, index = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=0)
        assert int((index >= 0).sum()) == ka[key]["covered"]
        assert [int((index == t).sum()) for t in range(vi.shape[0])] == ka[key]["per_triangle"]
        assert zlib.crc32(index.tobytes()) == ka[key]["index_crc32"]
    # 130 305 px for the README triangle, 103 240 px for the two-triangle demo
    assert ka["hello_triangle_512"]["covered"] == 130305
    assert ka["two_triangles_512"]["covered"] == 103240


def test_config3_checksum():
    from drtk_b200 import scenes
    with open(os.path.join(GOLDEN, "known_answers.json")) as f:
        ka = json.load(f)
    v, vi, H, W = scenes.config_mesh(3, N=1)
    _, index = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=0)
    assert int((index >= 0).sum()) == ka["config3_n1_1024"]["covered"]
    assert zlib.crc32(index.tobytes()) == ka["config3_n1_1024"]["index_crc32"]


def test_rasterize_edge_cases():
    # empty topology, degenerate / behind-camera / off-screen triangles -> empty image
    v = np.array([[[1, 1, 1], [5, 1, 1], [1, 5, 1], [1, 1, -1], [-9, -9, 1], [-5, -9, 1], [-9, -5, 1]]], np.float32)
    for vi in ([[0, 0, 0]], [[0, 1, 3]], [[4, 5, 6]], [[0, 1, 1]]):
        d, i = O.rasterize(v, np.array(vi, np.int32), 8, 8)
        assert (i == -1).all() and (d == 0).all()
    d, i = O.rasterize(v, np.zeros((0, 3), np.int32), 8, 8)
    assert (i == -1).all()
    # top nibble of vi[...,0] is ignored (reference :74)
    d0, i0 = O.rasterize(v, np.array([[0, 1, 2]], np.int32), 8, 8)
    d1, i1 = O.rasterize(v, np.array([[0 | (0x7 << 28), 1, 2]], np.int32), 8, 8)
    np.testing.assert_array_equal(i0, i1)
    assert (i0 >= 0).sum() > 0


# ------------------------------------------------------------------------------------------------
# wireframe mode: the fixtures wire_*.npz are outputs of the REFERENCE CUDA kernel (the reference has no
# CPU twin for this mode), written on a B200 by tests/golden/make_golden_wireframe.py
# ------------------------------------------------------------------------------------------------
def test_wireframe_oracle_against_reference_cuda_fixtures():
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, "wire_*.npz")))
    assert len(files) >= 3
    for fn in files:
        z = np.load(fn)
        d, i = O.rasterize_lines(z["v"], z["vi"], int(z["H"]), int(z["W"]))
        line_px = int((z["index_img"] >= 0).sum())
        assert line_px > 100
        # index_img: identical up to knife-edge pixels where MUFU.RCP (GPU) and 1/x (here) can round a crossing
        # point to different sides of a segment end -- none in these fixtures, and never more than 0.2 %
        bad = i != z["index_img"]
        assert int(bad.sum()) <= max(1, line_px // 500), (fn, int(bad.sum()))
        # depth: same formula, approximate vs exact reciprocals -> a few ulp
        ud = np.abs(d.view(np.int32).astype(np.int64) - z["depth_img"].view(np.int32).astype(np.int64))
        assert int(ud[~bad].max()) <= 8, (fn, int(ud[~bad].max()))
        # occluding interiors carry depth with index -1; the one-pixel canvas border is never written (:333-337)
        assert int(((z["index_img"] < 0) & (z["depth_img"] > 0)).sum()) > 0
        assert (i[:, 0, :] == -1).all() and (i[:, -1, :] == -1).all() and (i[:, :, 0] == -1).all() and (i[:, :, -1] == -1).all()


def test_wireframe_oracle_edge_flags():
    v = np.array([[[4, 4, 1], [28, 6, 1], [10, 26, 1]]], np.float32)
    base = np.array([[0, 1, 2]], np.int32)
    counts = []
    for flag in range(8):
        vi = base.copy(); vi[0, 0] |= flag << 28
        d, i = O.rasterize_lines(v, vi, 32, 32)
        counts.append(int((i >= 0).sum()))
        assert int((d > 0).sum()) > 100  # the interior occludes whatever the flags say
    assert counts[0] == 0 and counts[7] > counts[1] > 0 and counts[7] > counts[2] > 0 and counts[7] > counts[4] > 0
    assert counts[7] <= counts[1] + counts[2] + counts[4]  # shared corner pixels are counted once


# ------------------------------------------------------------------------------------------------
# sparse interpolation matrices: numpy oracle vs the reference's CPU implementation (tests/golden/mat_*.npz,
# written by tests/golden/make_golden_matrix.py from oracle/_ref)
# ------------------------------------------------------------------------------------------------
def _mat_files():
    import glob
    files = sorted(glob.glob(os.path.join(GOLDEN, "mat_*.npz")))
    assert len(files) >= 2
    return files


def test_interpolation_matrix_oracle_against_reference_cpu():
    for fn in _mat_files():
        z = np.load(fn)
        o = O.interpolation_matrix(z["vi"], z["index_img"], z["bary_img"], int(z["V"]))
        np.testing.assert_array_equal(o["row_pixels"], z["row_pixels"])
        np.testing.assert_array_equal(o["crow"], z["crow"])
        np.testing.assert_array_equal(o["col"], z["col"])
        np.testing.assert_array_equal(o["values"], z["values"])  # a pure permutation of the barycentrics
        assert (np.diff(o["col"].reshape(-1, 3), axis=1) > 0).all()  # ascending within a row
        # rows of A sum to 1 up to rounding (perspective-corrected barycentrics)
        np.testing.assert_allclose(o["dense"].sum(1), 1.0, atol=1e-5)
        # gradient of sum(values * w) w.r.t. bary: w scattered back through the same permutation
        N, H, W = z["index_img"].shape
        gb = np.zeros((N, 3, H * W), np.float32)
        n, hw = o["row_pixels"] // (H * W), o["row_pixels"] % (H * W)
        wv = z["w_values"].reshape(-1, 3)
        for k in range(3):
            gb[n, o["order"][:, k], hw] = wv[:, k]
        np.testing.assert_array_equal(gb.reshape(N, 3, H, W), z["grad_bary"])


def test_interpolation_normal_matrix_oracle_against_reference_cpu():
    for fn in _mat_files():
        z = np.load(fn)
        V = int(z["V"])
        o = O.interpolation_normal_matrix(z["vi"], z["index_img"], z["bary_img"], V)
        np.testing.assert_array_equal(o["crow"], z["n_crow"])
        np.testing.assert_array_equal(o["col"], z["n_col"])
        np.testing.assert_allclose(o["values"], z["n_values"], rtol=2e-5, atol=2e-5 * np.abs(z["n_values"]).max())
        # A^T A of the interpolation matrix, and symmetric
        A = O.interpolation_matrix(z["vi"], z["index_img"], z["bary_img"], V)["dense"]
        np.testing.assert_allclose(o["dense"], A.T @ A, rtol=1e-9, atol=1e-9)
        np.testing.assert_allclose(o["dense"], o["dense"].T, rtol=0, atol=1e-12)
