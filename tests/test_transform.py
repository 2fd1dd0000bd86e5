"""transform / project_points.

Golden vectors: tests/golden/transform_*.npz, written by tests/golden/make_golden_transform.py from the
UNMODIFIED reference `project_points` (float64, values + autograd gradients of a seeded linear loss).
CPU suite: the numpy oracle, the host's FOV estimators and the pure-torch statement against those vectors.
GPU suite (-m gpu): the CUDA kernels through the C ABI against the same vectors (fp32: rtol 1e-5 of scale) and,
on larger seeded clouds, against the oracle / the float64 torch statement.
"""
import glob
import os

import numpy as np
import pytest
import torch as th

import drtk_b200
import sys

T = sys.modules["drtk_b200.transform"]  # the module (the package attribute of that name is the function)
from oracle import oracle as O
from tests.util import GOLDEN, assert_close

NAMES = sorted(os.path.basename(p)[len("transform_"):-4] for p in glob.glob(os.path.join(GOLDEN, "transform_*.npz")))
PARAMS = ("v", "campos", "camrot", "focal", "princpt", "D")


def load(name):
    g = dict(np.load(os.path.join(GOLDEN, f"transform_{name}.npz")))
    mode = [None if m == "None" else str(m) for m in g.pop("mode")]
    return g, (mode if len(mode) > 1 else mode[0])


def test_fixture_set():
    assert set(NAMES) >= {"pinhole", "rt4", "rt5_fov", "rt8_fov", "fisheye", "fisheye_fov", "fisheye62", "fisheye62_fov", "mixed"}


# ---- CPU: oracle and host logic vs the reference's vectors ---------------------------------------------
@pytest.mark.parametrize("name", NAMES)
def test_oracle_forward_matches_reference(name):
    g, mode = load(name)
    vp, vc = O.transform_fwd(g["v"], g["campos"], g["camrot"], g["focal"], g["princpt"], mode, g.get("D"),
                             g.get("fov_used"), cull="fov" in g)
    assert_close(vc, g["v_cam"], rtol=1e-12, what="v_cam")
    assert_close(vp, g["v_pix"], rtol=1e-11, what="v_pix")


@pytest.mark.parametrize("name", NAMES)
def test_oracle_fd_gradients_match_reference(name):
    g, mode = load(name)
    fd = O.transform_vjp_fd(g["w_pix"], g["w_cam"], g["v"], g["campos"], g["camrot"], g["focal"], g["princpt"], mode,
                            g.get("D"), g.get("fov_used"), cull="fov" in g)
    for k in PARAMS:
        if k in fd:
            assert_close(fd[k], g["g_" + k], rtol=2e-5, what=f"grad {k}")


@pytest.mark.parametrize("name", [n for n in NAMES if not n.endswith("_fov") and n != "pinhole"])
def test_fov_estimators_match_reference(name):
    g, mode = load(name)
    D = th.from_numpy(g["D"])
    if isinstance(mode, list):
        for n, m in enumerate(mode):
            if m == "radial-tangential":
                assert_close(T.estimate_rt_fov(D[n:n + 1]).numpy(), g["fov_used"][n:n + 1], rtol=1e-6)
            elif m == "fisheye":
                assert_close(T.estimate_fisheye_fov(D[n:n + 1]).numpy(), g["fov_used"][n:n + 1], rtol=1e-6)
    else:
        est = T.estimate_rt_fov(D) if mode == "radial-tangential" else T.estimate_fisheye_fov(D)
        np.testing.assert_allclose(est.numpy(), g["fov_used"], rtol=1e-6)


def _torch_inputs(g, dev="cpu", dtype=th.float64):
    t = {k: th.from_numpy(g[k]).to(dev, dtype).requires_grad_(True) for k in PARAMS if k in g}
    fov = th.from_numpy(g["fov"]).to(dev, dtype) if "fov" in g else None
    return t, fov


def _run(fn, g, mode, dev="cpu", dtype=th.float64):
    t, fov = _torch_inputs(g, dev, dtype)
    v_pix, v_cam = fn(t["v"], t["campos"], t["camrot"], t["focal"], t["princpt"], mode, t.get("D"), fov)
    w_pix, w_cam = (th.from_numpy(g[k]).to(dev, dtype) for k in ("w_pix", "w_cam"))
    ((v_pix * w_pix).sum() + (v_cam * w_cam).sum()).backward()
    grads = {k: (x.grad if x.grad is not None else th.zeros_like(x)) for k, x in t.items()}
    return v_pix.detach(), v_cam.detach(), grads


@pytest.mark.parametrize("name", NAMES)
def test_torch_statement_matches_reference(name):
    g, mode = load(name)
    v_pix, v_cam, grads = _run(T.project_points_ref, g, mode)
    assert_close(v_pix.numpy(), g["v_pix"], rtol=1e-11, what="v_pix")
    assert_close(v_cam.numpy(), g["v_cam"], rtol=1e-12, what="v_cam")
    for k, x in grads.items():
        assert_close(x.numpy(), g["g_" + k], rtol=1e-9, what=f"grad {k}")


def test_hello_triangle_cpu():
    """BASELINE config 1: README triangle through transform with an identity camera, CPU tensors."""
    v = th.tensor([[[0.0, 511.0, 1.0], [255.0, 0.0, 1.0], [511.0, 511.0, 1.0]]])
    out = drtk_b200.transform(v, campos=th.zeros(1, 3), camrot=th.eye(3)[None], focal=th.eye(2)[None], princpt=th.zeros(1, 2))
    assert th.equal(out, v)


def test_argument_errors():
    v = th.zeros(1, 4, 3)
    with pytest.raises(ValueError, match="exactly one of Rt"):
        drtk_b200.transform(v, focal=th.eye(2)[None], princpt=th.zeros(1, 2))
    with pytest.raises(ValueError, match="exactly one of K"):
        drtk_b200.transform(v, campos=th.zeros(1, 3), camrot=th.eye(3)[None])
    with pytest.raises(ValueError, match="Invalid distortion mode"):
        drtk_b200.transform(v, campos=th.zeros(1, 3), camrot=th.eye(3)[None], focal=th.eye(2)[None], princpt=th.zeros(1, 2),
                            distortion_mode="barrel", distortion_coeff=th.zeros(1, 4))
    with pytest.raises(ValueError, match="Invalid distortion mode"):
        drtk_b200.transform(th.zeros(2, 4, 3), campos=th.zeros(2, 3), camrot=th.eye(3)[None].expand(2, -1, -1),
                            focal=th.eye(2)[None].expand(2, -1, -1), princpt=th.zeros(2, 2),
                            distortion_mode=["pinhole", "fisheye62"], distortion_coeff=th.zeros(2, 8))


def test_rt_and_K_forms_cpu():
    g, _ = load("pinhole")
    t, _ = _torch_inputs(g)
    R, c = t["camrot"].detach(), t["campos"].detach()
    Rt = th.cat((R, -(R @ c[..., None])), -1)
    K = th.zeros(R.shape[0], 3, 3, dtype=R.dtype)
    K[:, :2, :2] = t["focal"].detach(); K[:, :2, 2] = t["princpt"].detach(); K[:, 2, 2] = 1
    out = drtk_b200.transform(t["v"].detach(), K=K, Rt=Rt)
    assert_close(out.numpy(), g["v_pix"], rtol=1e-9)


# ---- GPU: the CUDA kernels through the C ABI --------------------------------------------------------
DEV = "cuda:0"


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_matches_reference_vectors(name):
    g, mode = load(name)
    v_pix, v_cam, grads = _run(T.project_points, g, mode, DEV, th.float32)
    assert_close(v_cam.cpu().numpy(), g["v_cam"], rtol=1e-5, what="v_cam")
    assert_close(v_pix.cpu().numpy(), g["v_pix"], rtol=2e-5, what="v_pix")
    for k, x in grads.items():
        assert_close(x.cpu().numpy(), g["g_" + k], rtol=5e-5, what=f"grad {k}")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,nd", [(None, 0), ("radial-tangential", 8), ("fisheye", 4), ("fisheye62", 8)])
def test_cuda_large_cloud_vs_oracle_and_fp64(mode, nd):
    """50k vertices x 4 cameras (the per-camera gradient is a 50k-term reduction)."""
    N, V = 4, 50625
    gen = th.Generator().manual_seed(91)
    from tests.util import random_cameras as cameras
    campos, camrot, focal, princpt = cameras(N, gen)
    vbig = th.rand((N, V, 4), generator=gen, dtype=th.float64) * th.tensor([2.0, 2.0, 1.0, 1.0], dtype=th.float64) + th.tensor([-1.0, -1.0, 2.0, 0.0], dtype=th.float64)
    v = vbig[..., :3]  # row stride 4: not contiguous
    D = (th.rand((N, nd), generator=gen, dtype=th.float64) * 2 - 1) * 0.02 if nd else None
    fov = th.full((N, 1), 0.45, dtype=th.float64) if mode else None
    w = th.rand((N, V, 3), generator=gen, dtype=th.float64)

    def run(fn, dev, dt):
        leaves = [x.detach().to(dev, dt).clone().requires_grad_(True) for x in (v, campos, camrot, focal, princpt)]
        Dl = D.detach().to(dev, dt).clone().requires_grad_(True) if D is not None else None
        vp, vc = fn(*leaves, mode, Dl, None if fov is None else fov.to(dev, dt))
        (vp * w.to(dev, dt)).sum().backward()
        return vp.detach().cpu().numpy(), vc.detach().cpu().numpy(), [x.grad.cpu().numpy() for x in leaves + ([Dl] if nd else [])]

    vp, vc, gr = run(T.project_points, DEV, th.float32)
    vp64, vc64, gr64 = run(T.project_points_ref, "cpu", th.float64)
    vpo, vco = O.transform_fwd(v.numpy(), campos.numpy(), camrot.numpy(), focal.numpy(), princpt.numpy(), mode,
                               None if D is None else D.numpy(), None if fov is None else fov.numpy(), cull=mode == "fisheye62")
    assert_close(vp64, vpo, rtol=1e-11, what="torch statement vs oracle")
    assert_close(vp, vpo, rtol=1e-5, what="v_pix vs oracle")
    assert_close(vc, vco, rtol=1e-5, what="v_cam vs oracle")
    for a, e, k in zip(gr, gr64, PARAMS):
        assert_close(a, e, rtol=5e-5, what=f"grad {k}")


@pytest.mark.gpu
def test_cuda_lut_and_transform_wrappers():
    g, mode = load("fisheye62_fov")
    t, fov = _torch_inputs(g, DEV, th.float32)
    gen = th.Generator().manual_seed(5)
    lut = (th.rand((3, 2, 9, 11), generator=gen) - 0.5).to(DEV)
    spacing = th.tensor([[80.0, 64.0]]).expand(3, -1).to(DEV)
    args = [t[k].detach() for k in ("v", "campos", "camrot", "focal", "princpt")]
    a, _ = T.project_points(*args, "fisheye62_lut", t["D"].detach(), fov, lut, spacing)
    b, _ = T.project_points_ref(*[x.double() for x in args], "fisheye62_lut", t["D"].detach().double(), fov.double(),
                                lut.double(), spacing.double())
    assert_close(a.cpu().numpy(), b.cpu().numpy(), rtol=2e-5, what="fisheye62_lut")
    # Rt / K form, autocast-style half input, and the package-level names
    R, c = args[2], args[1]
    Rt = th.cat((R, -(R @ c[..., None])), -1)
    K = th.zeros(3, 3, 3, device=DEV)
    K[:, :2, :2] = args[3]; K[:, :2, 2] = args[4]; K[:, 2, 2] = 1
    out = drtk_b200.transform(args[0], K=K, Rt=Rt)
    gp, _ = load("fisheye62_fov")[0], None
    ref, _ = T.project_points_ref(*[x.double() for x in args])
    assert_close(out.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, what="K/Rt pinhole")
    vp, vc = drtk_b200.transform_with_v_cam(args[0], *args[1:])
    assert vp.shape == vc.shape == args[0].shape
    # strided vertices (row stride 4) and an expanded camera give the same bits as dense copies
    vb = th.zeros(3, args[0].shape[1], 4, device=DEV)
    vb[..., :3] = args[0]
    vs = vb[..., :3].requires_grad_(True)
    assert not vs.is_contiguous()
    vp2, _ = drtk_b200.transform_with_v_cam(vs, *args[1:])
    assert th.equal(vp2, vp)
    vp2.sum().backward()
    vd = args[0].clone().requires_grad_(True)
    drtk_b200.transform(vd, *args[1:]).sum().backward()
    assert th.equal(vs.grad, vd.grad)
