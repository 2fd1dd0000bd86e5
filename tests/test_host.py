"""CPU suite: host-side logic of the drop-in package (no GPU, no kernel launches)."""
import os
import sys

import pytest
import torch as th

import drtk_b200
from drtk_b200 import dist as ddist
from drtk_b200 import scenes


def test_public_surface_matches_reference_names():
    # drtk/__init__.py:8-33 (hot-path subset)
    for name in ("rasterize", "rasterize_with_depth", "render", "interpolate", "edge_grad_estimator",
                 "transform", "transform_with_v_cam"):
        assert callable(getattr(drtk_b200, name))


def test_install_as_drtk_shim():
    drtk_b200.install_as_drtk()
    import drtk
    from drtk.render import render
    from drtk.edge_grad_estimator import edge_grad_estimator
    assert drtk is drtk_b200 and render is drtk_b200.render
    assert edge_grad_estimator is drtk_b200.edge_grad_estimator
    del sys.modules["drtk"]


def test_transform_hello_triangle_cpu():
    """BASELINE config 1: the README triangle through an identity pinhole camera on CPU tensors."""
    v, vi, H, W = scenes.hello_triangle()
    out = drtk_b200.transform(v, campos=th.zeros(1, 3), camrot=th.eye(3)[None], focal=th.eye(2)[None],
                              princpt=th.zeros(1, 2))
    assert th.equal(out, v)


def test_transform_matches_manual_projection_and_Rt_K_forms():
    g = th.Generator().manual_seed(0)
    N, V = 2, 50
    v = th.randn(N, V, 3, generator=g, dtype=th.float64)
    A = th.randn(N, 3, 3, generator=g, dtype=th.float64)
    R, _ = th.linalg.qr(A)
    c = th.randn(N, 3, generator=g, dtype=th.float64) - th.tensor([0, 0, 6.0], dtype=th.float64)
    f = th.tensor([[500.0, 2.0], [0.0, 480.0]], dtype=th.float64).expand(N, 2, 2)
    pp = th.tensor([256.0, 250.0], dtype=th.float64).expand(N, 2)
    out = drtk_b200.transform(v, campos=c, camrot=R, focal=f, princpt=pp)
    vc = (R[:, None] @ (v - c[:, None])[..., None])[..., 0]
    xy = (f[:, None] @ (vc[..., :2] / vc[..., 2:3])[..., None])[..., 0] + pp[:, None]
    assert th.allclose(out[..., :2], xy, rtol=1e-12, atol=1e-9) and th.allclose(out[..., 2], vc[..., 2])
    t = -(R @ c[..., None])
    K = th.zeros(N, 3, 3, dtype=th.float64)
    K[:, :2, :2] = f; K[:, :2, 2] = pp; K[:, 2, 2] = 1
    out2 = drtk_b200.transform(v, K=K, Rt=th.cat((R, t), -1))
    assert th.allclose(out, out2, rtol=1e-10, atol=1e-8)
    with pytest.raises(ValueError):
        drtk_b200.transform(v, campos=c, camrot=R, focal=f, princpt=pp, K=K)
    # zero distortion coefficients: the fisheye model reduces to theta/r scaling of the pinhole image plane
    fe = drtk_b200.transform(v, campos=c, camrot=R, focal=f, princpt=pp, distortion_mode="fisheye",
                             distortion_coeff=th.zeros(N, 4, dtype=th.float64))
    assert fe.shape == out.shape and th.allclose(fe[..., 2], out[..., 2])


def test_transform_is_differentiable():
    v = th.randn(1, 4, 3, dtype=th.float64) + th.tensor([0, 0, 5.0], dtype=th.float64)
    v.requires_grad_(True)
    args = dict(campos=th.zeros(1, 3, dtype=th.float64), camrot=th.eye(3, dtype=th.float64)[None],
                focal=th.eye(2, dtype=th.float64)[None] * 100, princpt=th.zeros(1, 2, dtype=th.float64))
    assert th.autograd.gradcheck(lambda x: drtk_b200.transform(x, **args), (v,))


def test_argument_checks_raise_like_the_reference():
    v = th.zeros(1, 3, 3)
    vi = th.zeros(1, 1, 3, dtype=th.int32)
    with pytest.raises(RuntimeError, match=r"rasterize\(\): expected all inputs to be on same cuda device"):
        drtk_b200.rasterize(v, vi, 8, 8)
    idx = th.zeros(1, 8, 8, dtype=th.int32)
    with pytest.raises(RuntimeError, match=r"render\(\): expected all inputs to be on same cuda device"):
        drtk_b200.render(v, vi, idx)
    with pytest.raises(RuntimeError, match=r"interpolate\(\)"):
        drtk_b200.interpolate(v, vi, idx, th.zeros(1, 3, 8, 8))
    with pytest.raises(RuntimeError, match=r"edge_grad_estimator\(\)"):
        drtk_b200.edge_grad_estimator(v, vi, th.zeros(1, 3, 8, 8), th.zeros(1, 2, 8, 8), idx)


def test_missing_native_library_fails_loudly(monkeypatch):
    from drtk_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdrtk_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _lib.load()


def test_scenes_are_deterministic_and_sized_like_baseline():
    v1, vi1 = scenes.grid_mesh(11, 9, 64, 80, 2, seed=4)
    v2, vi2 = scenes.grid_mesh(11, 9, 64, 80, 2, seed=4)
    assert th.equal(v1, v2) and th.equal(vi1, vi2)
    assert vi1.shape == (2 * 10 * 8, 3) and v1.shape == (2, 99, 3) and vi1.dtype == th.int32
    assert not th.equal(v1[0], v1[1])
    for cfg, F, V in ((3, 5000, 2601), (4, 100352, 50625)):
        c = scenes.CONFIGS[cfg]
        assert 2 * (c["nx"] - 1) * (c["ny"] - 1) == F and c["nx"] * c["ny"] == V
    assert scenes.grid_mesh(5, 5, 32, 32, 1, seed=0, overdraw=True)[1].shape == (64, 3)


def test_shard_batch_partitions():
    for n, w in ((64, 8), (10, 4), (3, 8), (0, 2)):
        spans = [ddist.shard_batch(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [e - b for b, e in spans]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        ddist.shard_batch(8, 8, 8)


def test_allreduce_shared_grads_single_process():
    a, b = th.arange(12.0).view(2, 2, 3), th.ones(2, 2, 4)
    (ra, rb), work = ddist.allreduce_shared_grads([a, None, b])
    assert work is None and th.equal(ra, a.sum(0)) and th.equal(rb, b.sum(0))


def test_pure_torch_refs_match_the_oracle():
    """render_ref / interpolate_ref (float64 PyTorch statements of the ops, any device) against the C oracle."""
    import numpy as np
    import torch as th

    import drtk_b200
    from drtk_b200 import scenes
    from oracle import oracle as O

    H, W = 40, 56
    v, vi = scenes.grid_mesh(7, 6, H, W, 2, seed=3, overdraw=True)
    _, index = O.rasterize(v.numpy(), vi.numpy(), H, W, mode=0)
    d_o, b_o = O.render_fwd(v.double().numpy(), vi.numpy(), index)
    d, b = drtk_b200.render_ref(v.double(), vi, th.from_numpy(index))
    np.testing.assert_allclose(d.numpy(), d_o, rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(b.numpy(), b_o, rtol=1e-9, atol=1e-9)
    attr = scenes.vertex_attributes(2, v.shape[1], 5, seed=4)
    img_o = O.interpolate_fwd(attr.double().numpy(), vi.numpy(), index, b_o)
    img = drtk_b200.interpolate_ref(attr.double(), vi, th.from_numpy(index), th.from_numpy(b_o))
    np.testing.assert_allclose(img.numpy(), img_o, rtol=1e-6, atol=1e-7)  # the oracle evaluates the empty-pixel sweep in float
    assert (index == -1).any()  # the coordinate sweep of empty pixels is part of the comparison
    # differentiable: gradients w.r.t. vertices flow through render_ref
    vv = v.double().requires_grad_(True)
    drtk_b200.render_ref(vv, vi, th.from_numpy(index))[1].sum().backward()
    assert vv.grad is not None and bool(th.isfinite(vv.grad).all())


def test_product_never_touches_the_oracle_or_the_reference():
    """The oracle / oracle/_ref are test infrastructure: no module of the product may import or open them, and every
    op with a kernel in the reference refuses CPU tensors instead of falling back."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for f in glob.glob(os.path.join(root, "drtk_b200", "*.py")):
        src = open(f).read()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f
        assert "oracle/" not in src and "_ref/" not in src and "/root/reference" not in src, f
    x = th.zeros(1, 3, 3)
    vi = th.zeros(1, 3, dtype=th.int32)
    idx = th.zeros(1, 4, 4, dtype=th.int32)
    bary = th.zeros(1, 3, 4, 4)
    for call in (lambda: drtk_b200.rasterize(x, vi, 4, 4), lambda: drtk_b200.render(x, vi, idx),
                 lambda: drtk_b200.interpolate(x, vi, idx, bary),
                 lambda: drtk_b200.edge_grad_estimator(x, vi, bary, bary, idx).sum().backward(),
                 lambda: drtk_b200.grid_scatter(bary, th.zeros(1, 4, 4, 2), 4, 4),
                 lambda: drtk_b200.mipmap_grid_sample([bary], th.zeros(1, 4, 4, 2), th.zeros(1, 4, 4, 2, 2), 1)):
        with pytest.raises(RuntimeError):
            call()


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver launches first) prints exactly one JSON line with the
    contract's keys; it needs no GPU."""
    import json
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "impl", "cpu_baseline", "e2e", "gpu_launches"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "Mpix/s" and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "transform_only" in d["cpu_baseline"]
