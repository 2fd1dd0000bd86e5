"""mipmap_grid_sample / grid_scatter.

Fixtures tests/golden/samp_*.npz come from the UNMODIFIED reference's pure-PyTorch statements (`grid_scatter_ref`,
`mipmap_grid_sample_ref`; generator tests/golden/make_golden_samplers.py), float64.
CPU suite: the numpy oracle (oracle/samplers.py) and this package's own torch statements against those vectors.
GPU suite (-m gpu): the CUDA kernels through the C ABI against the fixtures, against the oracle on seeded inputs that
exercise what the torch statements cannot (per-pixel sample counts, clip_grad, align_corners, out-of-image bicubic
splats) and against the reference's CUDA kernels from oracle/_ref when they travelled to the box.
fp32 tolerance: 1e-5 of the tensor's scale (2e-5 for accumulated gradients).
"""
import glob
import os

import numpy as np
import pytest
import torch as th

import drtk_b200
from oracle import ref as R
from oracle import samplers as S
from tests.util import GOLDEN, assert_close

SCATTER = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "samp_scatter_*.npz")))
MIPMAP = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "samp_mipmap_*.npz")))
MODE = {0: "bilinear", 2: "bicubic"}
PAD = {0: "zeros", 1: "border", 2: "reflection"}
DEV = "cuda:0"


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def levels_of(g):
    return [g[f"level{l}"] for l in range(int(g["meta"][3]))]


def test_fixture_sets():
    assert len(SCATTER) == 12 and len(MIPMAP) == 24 and len(CUDA_MIP) == 5 and len(CUDA_SCATTER) == 4


# ---- CPU: oracle vs the reference's statements -------------------------------------------------------
@pytest.mark.parametrize("name", SCATTER)
def test_oracle_grid_scatter(name):
    g = load(name)
    Ho, Wo, interp, pad, align = map(int, g["meta"])
    assert_close(S.grid_scatter_fwd(g["input"], g["grid"], Ho, Wo, interp, pad, bool(align)), g["out"], rtol=1e-11, what="out")
    gi, gg = S.grid_scatter_bwd(g["w"], g["input"], g["grid"], interp, pad, bool(align))
    assert_close(gi, g["g_input"], rtol=1e-11, what="grad input")
    assert_close(gg, g["g_grid"], rtol=1e-10, what="grad grid")


@pytest.mark.parametrize("name", MIPMAP)
def test_oracle_mipmap_grid_sample(name):
    g = load(name)
    aniso, interp, pad, _ = map(int, g["meta"])
    lv = levels_of(g)
    out = S.mipmap_grid_sample_fwd(lv, g["grid"], g["jac"], aniso, interp, pad, False, force_max_aniso=True)
    assert_close(out, g["out"], rtol=1e-10, what="out")
    gl, gg = S.mipmap_grid_sample_bwd(g["w"], lv, g["grid"], g["jac"], aniso, interp, pad, False, force_max_aniso=True)
    for l, x in enumerate(gl):
        assert_close(x, g[f"g_level{l}"], rtol=1e-10, what=f"grad level {l}")
    assert_close(gg, g["g_grid"], rtol=1e-9, what="grad grid")


@pytest.mark.parametrize("name", SCATTER[::3])
def test_torch_statement_grid_scatter(name):
    g = load(name)
    Ho, Wo, interp, pad, align = map(int, g["meta"])
    x, grid = th.from_numpy(g["input"]).requires_grad_(True), th.from_numpy(g["grid"]).requires_grad_(True)
    out = drtk_b200.grid_scatter_ref(x, grid, Ho, Wo, MODE[interp], PAD[pad], bool(align))
    (out * th.from_numpy(g["w"])).sum().backward()
    assert_close(out.detach().numpy(), g["out"], rtol=1e-12)
    assert_close(x.grad.numpy(), g["g_input"], rtol=1e-12)
    assert_close(grid.grad.numpy(), g["g_grid"], rtol=1e-12)


@pytest.mark.parametrize("name", MIPMAP[::3])
def test_torch_statement_mipmap(name):
    g = load(name)
    aniso, interp, pad, _ = map(int, g["meta"])
    lv = [th.from_numpy(x).requires_grad_(True) for x in levels_of(g)]
    grid = th.from_numpy(g["grid"]).requires_grad_(True)
    out = drtk_b200.mipmap_grid_sample_ref(lv, grid, th.from_numpy(g["jac"]), aniso, MODE[interp], PAD[pad], False)
    (out * th.from_numpy(g["w"])).sum().backward()
    assert_close(out.detach().numpy(), g["out"], rtol=1e-11)
    assert_close(grid.grad.numpy(), g["g_grid"], rtol=1e-10)
    for l, x in enumerate(lv):
        assert_close(x.grad.numpy(), g[f"g_level{l}"], rtol=1e-11)


# ---- CPU: oracle vs outputs of the reference CUDA kernels (captured on a B200 by make_golden_samplers_cuda.py) -----
CUDA_MIP = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "sampcuda_mipmap_*.npz")))
CUDA_SCATTER = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "sampcuda_scatter_*.npz")))


def _frac_bad(a, e, rtol):
    a, e = np.asarray(a, np.float64), np.asarray(e, np.float64)
    return float((np.abs(a - e) > rtol * np.abs(e) + rtol * np.abs(e).max()).mean())


@pytest.mark.parametrize("name", CUDA_MIP)
def test_oracle_mipmap_vs_reference_cuda_vectors(name):
    """Per-pixel sample counts, clip_grad, the forward's align_corners override, non-square textures.  The fixtures are
    fp32 GPU results; a pixel whose footprint sits on a level / sample-count boundary may fall on the other side of it."""
    g = load(name)
    aniso, interp, pad, nlev, align, force, clip = map(int, g["meta"])
    lv = levels_of(g)
    out = S.mipmap_grid_sample_fwd(lv, g["grid"], g["jac"], aniso, interp, pad, bool(align), bool(force), bool(clip))
    assert _frac_bad(out, g["out"], 1e-5) < 5e-3, "out"
    gl, gg = S.mipmap_grid_sample_bwd(g["w"], lv, g["grid"], g["jac"], aniso, interp, pad, bool(align), bool(force), bool(clip))
    assert _frac_bad(gg, g["g_grid"], 3e-5) < 5e-3, "grad grid"
    for l, x in enumerate(gl):
        assert _frac_bad(x, g[f"g_level{l}"], 3e-5) < 2e-2, f"grad level {l}"


@pytest.mark.parametrize("name", CUDA_SCATTER)
def test_oracle_grid_scatter_vs_reference_cuda_vectors(name):
    """Splat positions far outside the image: the bicubic kernel pads the sample position itself."""
    g = load(name)
    Ho, Wo, interp, pad, align = map(int, g["meta"])
    assert _frac_bad(S.grid_scatter_fwd(g["input"], g["grid"], Ho, Wo, interp, pad, bool(align)), g["out"], 1e-5) < 2e-3
    gi, gg = S.grid_scatter_bwd(g["w"], g["input"], g["grid"], interp, pad, bool(align))
    assert _frac_bad(gi, g["g_input"], 1e-5) < 2e-3 and _frac_bad(gg, g["g_grid"], 3e-5) < 2e-3


def test_argument_errors_cpu():
    x, grid = th.zeros(1, 2, 4, 4), th.zeros(1, 4, 4, 2)
    with pytest.raises(ValueError, match="only 'bilinear' and 'bicubic'"):
        drtk_b200.grid_scatter(x, grid, 4, 4, mode="nearest")
    with pytest.raises(ValueError, match="padding_mode"):
        drtk_b200.mipmap_grid_sample([x], grid, th.zeros(1, 4, 4, 2, 2), 1, padding_mode="wrap")
    with pytest.raises(RuntimeError, match="same device"):  # CPU tensors: no fallback
        drtk_b200.grid_scatter(x, grid, 4, 4)
    with pytest.raises(RuntimeError, match="same device"):
        drtk_b200.mipmap_grid_sample([x], grid, th.zeros(1, 4, 4, 2, 2), 1)


# ---- GPU ---------------------------------------------------------------------------------------------
def cu(a, grad=False):
    return th.as_tensor(a).to(DEV, th.float32).requires_grad_(grad)


@pytest.mark.gpu
@pytest.mark.parametrize("name", SCATTER)
def test_cuda_grid_scatter_vs_reference_vectors(name):
    g = load(name)
    Ho, Wo, interp, pad, align = map(int, g["meta"])
    x, grid = cu(g["input"], True), cu(g["grid"], True)
    out = drtk_b200.grid_scatter(x, grid, Ho, Wo, MODE[interp], PAD[pad], bool(align))
    (out * cu(g["w"])).sum().backward()
    assert_close(out.detach().cpu().numpy(), g["out"], rtol=1e-5, what="out")
    assert_close(x.grad.cpu().numpy(), g["g_input"], rtol=1e-5, what="grad input")
    assert_close(grid.grad.cpu().numpy(), g["g_grid"], rtol=2e-5, what="grad grid")


@pytest.mark.gpu
@pytest.mark.parametrize("name", MIPMAP)
def test_cuda_mipmap_vs_reference_vectors(name):
    g = load(name)
    aniso, interp, pad, _ = map(int, g["meta"])
    lv = [cu(x, True) for x in levels_of(g)]
    grid = cu(g["grid"], True)
    out = drtk_b200.mipmap_grid_sample(lv, grid, cu(g["jac"]), aniso, MODE[interp], PAD[pad], False, force_max_aniso=True)
    (out * cu(g["w"])).sum().backward()
    assert_close(out.detach().cpu().numpy(), g["out"], rtol=1e-5, what="out")
    assert_close(grid.grad.cpu().numpy(), g["g_grid"], rtol=2e-5, what="grad grid")
    for l, x in enumerate(lv):
        assert_close(x.grad.cpu().numpy(), g[f"g_level{l}"], rtol=2e-5, what=f"grad level {l}")


def _mip_inputs(seed, N=2, C=5, H=37, W=53, S=(64, 48), nlev=4, reach=1.2):
    g = th.Generator().manual_seed(seed)
    levels = [th.rand((N, C, max(S[0] >> l, 1), max(S[1] >> l, 1)), generator=g) for l in range(nlev)]
    grid = (th.rand((N, H, W, 2), generator=g) * 2 - 1) * reach
    jac = (th.rand((N, H, W, 2, 2), generator=g) - 0.5) * th.tensor([0.4, 0.04])[:, None] * th.rand((N, H, W, 1, 1), generator=g)
    w = th.rand((N, C, H, W), generator=g)
    return levels, grid, jac, w


def _run_mip(fn, levels, grid, jac, w, *args, **kw):
    lv = [t.clone().to(DEV).requires_grad_(True) for t in levels]
    gr = grid.clone().to(DEV).requires_grad_(True)
    out = fn(lv, gr, jac.to(DEV), *args, **kw)
    (out * w.to(DEV)).sum().backward()
    return [out.detach().cpu().numpy(), gr.grad.cpu().numpy()] + [t.grad.cpu().numpy() for t in lv]


MIP_VARIANTS = [
    # (mode, pad, max_aniso, align, force, clip, levels)
    ("bilinear", "zeros", 8, False, False, False, 4),
    ("bilinear", "border", 4, True, False, True, 2),
    ("bilinear", "reflection", 2, False, True, True, 3),
    ("bicubic", "zeros", 4, True, False, False, 3),
    ("bicubic", "border", 3, False, False, True, 1),
    ("bicubic", "reflection", 5, True, True, False, 4),
]


# bicubic + reflection + align_corners=True: the reflected tap coordinates are integers, so `floor(c / span)` sits
# exactly on a flip boundary; the reference's --use_fast_math division (x * rcp(span)) lands on either side depending
# on the texture size.  drtk_b200 reproduces the reference CUDA build there (same expression, same flags: see
# test_cuda_mipmap_vs_reference_cuda); the IEEE oracle cannot, so it checks that combination with align_corners=False.
ORACLE_VARIANTS = [v if v[:2] != ("bicubic", "reflection") else v[:3] + (False,) + v[4:] for v in MIP_VARIANTS]


@pytest.mark.gpu
@pytest.mark.parametrize("mode,pad,aniso,align,force,clip,nlev", ORACLE_VARIANTS)
def test_cuda_mipmap_vs_oracle(mode, pad, aniso, align, force, clip, nlev):
    """Non-square textures, per-pixel sample counts, clip_grad, align_corners (ignored forward, honoured backward).
    Level / sample-count selection is discrete: a pixel whose footprint sits on a decision boundary may legitimately
    fall on the other side in fp32, hence the (tiny) allowed fraction of outliers."""
    levels, grid, jac, w = _mip_inputs(31 + aniso, nlev=nlev)
    got = _run_mip(drtk_b200.mipmap_grid_sample, levels, grid, jac, w, aniso, mode, pad, align, force, clip)
    interp, p = {"bilinear": 0, "bicubic": 2}[mode], {"zeros": 0, "border": 1, "reflection": 2}[pad]
    lv = [t.numpy() for t in levels]
    out = S.mipmap_grid_sample_fwd(lv, grid.numpy(), jac.numpy(), aniso, interp, p, align, force, clip)
    gl, gg = S.mipmap_grid_sample_bwd(w.numpy(), lv, grid.numpy(), jac.numpy(), aniso, interp, p, align, force, clip)
    assert _frac_bad(got[0], out, 1e-5) < 2e-3, "out"
    assert _frac_bad(got[1], gg, 3e-5) < 2e-3, "grad grid"
    for a, e in zip(got[2:], gl):
        assert _frac_bad(a, e, 3e-5) < 5e-3, "grad level"


@pytest.mark.gpu
@pytest.mark.skipif(not R.samplers_available(), reason="oracle/_ref sampler extensions not present")
@pytest.mark.parametrize("mode,pad,aniso,align,force,clip,nlev", MIP_VARIANTS)
def test_cuda_mipmap_vs_reference_cuda(mode, pad, aniso, align, force, clip, nlev):
    levels, grid, jac, w = _mip_inputs(77 + aniso, N=2, C=4, H=96, W=128, S=(128, 96), nlev=nlev)
    got = _run_mip(drtk_b200.mipmap_grid_sample, levels, grid, jac, w, aniso, mode, pad, align, force, clip)
    ref = _run_mip(R.mipmap_grid_sample, levels, grid, jac, w, aniso, mode, pad, align, force, clip)
    assert _frac_bad(got[0], ref[0], 1e-5) < 1e-3, "out"
    assert _frac_bad(got[1], ref[1], 3e-5) < 1e-3, "grad grid"
    for a, e in zip(got[2:], ref[2:]):
        assert _frac_bad(a, e, 3e-5) < 2e-3, "grad level"


SCATTER_VARIANTS = [(m, p, a) for m in ("bilinear", "bicubic") for p in ("zeros", "border", "reflection") for a in (False, True)]


def _run_scatter(fn, x, grid, w, Ho, Wo, mode, pad, align):
    xl, gl = x.clone().to(DEV).requires_grad_(True), grid.clone().to(DEV).requires_grad_(True)
    out = fn(xl, gl, Ho, Wo, mode, pad, align)
    (out * w.to(DEV)).sum().backward()
    return out.detach().cpu().numpy(), xl.grad.cpu().numpy(), gl.grad.cpu().numpy()


@pytest.mark.gpu
@pytest.mark.parametrize("mode,pad,align", SCATTER_VARIANTS)
def test_cuda_grid_scatter_vs_oracle_and_reference_cuda(mode, pad, align):
    """Splat positions reach well outside the image (the bicubic centre-padding quirk), strided input."""
    g = th.Generator().manual_seed(5)
    N, C, H, W, Ho, Wo = 2, 5, 41, 35, 30, 44
    x = th.rand((N, H, W, C), generator=g).permute(0, 3, 1, 2)  # channels-last strides
    grid = (th.rand((N, H, W, 2), generator=g) * 2 - 1) * 1.5
    w = th.rand((N, C, Ho, Wo), generator=g)
    got = _run_scatter(drtk_b200.grid_scatter, x, grid, w, Ho, Wo, mode, pad, align)
    interp, p = {"bilinear": 0, "bicubic": 2}[mode], {"zeros": 0, "border": 1, "reflection": 2}[pad]
    out = S.grid_scatter_fwd(x.numpy(), grid.numpy(), Ho, Wo, interp, p, align)
    gi, gg = S.grid_scatter_bwd(w.numpy(), x.numpy(), grid.numpy(), interp, p, align)
    for a, e, what in zip(got, (out, gi, gg), ("out", "grad input", "grad grid")):
        assert _frac_bad(a, e, 2e-5) < 1e-3, what
    if R.samplers_available():
        ref = _run_scatter(R.grid_scatter, x, grid, w, Ho, Wo, mode, pad, align)
        for a, e, what in zip(got, ref, ("out", "grad input", "grad grid")):
            assert _frac_bad(a, e, 2e-5) < 1e-3, what + " vs reference CUDA"


@pytest.mark.gpu
def test_cuda_sampler_edge_cases():
    # empty batch / no gradient requested / single level / needs-grad gating
    lv = [th.rand(1, 2, 8, 8, device=DEV)]
    grid = th.zeros(1, 3, 3, 2, device=DEV)
    jac = th.zeros(1, 3, 3, 2, 2, device=DEV)  # zero footprint: p = 1e-6, N = 1, level 0
    out = drtk_b200.mipmap_grid_sample(lv, grid, jac, 4)
    ref = th.nn.functional.grid_sample(lv[0], grid, align_corners=False)
    assert th.allclose(out, ref, atol=1e-6) and not out.requires_grad
    assert drtk_b200.mipmap_grid_sample([th.rand(0, 2, 8, 8, device=DEV)], grid[:0], jac[:0], 2).shape == (0, 2, 3, 3)
    x = th.rand(1, 2, 3, 3, device=DEV, requires_grad=True)
    o = drtk_b200.grid_scatter(x, grid, 5, 5)
    o.sum().backward()
    assert x.grad.shape == x.shape and th.allclose(o.sum(), x.sum(), rtol=1e-5)  # border padding conserves mass
    with pytest.raises(RuntimeError, match="float32 only"):
        drtk_b200.grid_scatter(x.double(), grid.double(), 5, 5)
