"""Shared comparison helpers for the test-suite."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    # wire_*.npz (wireframe, reference CUDA outputs) and mat_*.npz (interpolation matrices, reference CPU outputs)
    # and transform_*.npz (projection, reference project_points outputs) are handled by their own tests
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith(("wire_", "mat_", "transform_", "samp_", "sampcuda_", "uvd_")))


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def ulp_diff(a, b):
    """|a - b| in units of float32 bit patterns (valid for same-sign finite floats)."""
    a = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.int64)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32).astype(np.int64)
    return np.abs(a - b)


def assert_close(actual, expected, rtol=1e-5, scale_rtol=None, what=""):
    """Element-wise |a-e| <= rtol*|e| + atol, with atol = (scale_rtol or rtol) * max|e|.

    Relative tolerance 1e-5 is the north-star bar for fp32 outputs; the absolute floor is the
    same fraction of the tensor's scale (a pure relative test is meaningless next to zeros,
    e.g. barycentrics on an edge)."""
    a = np.asarray(actual, np.float64)
    e = np.asarray(expected, np.float64)
    assert a.shape == e.shape, f"{what}: shape {a.shape} vs {e.shape}"
    scale = float(np.abs(e).max()) if e.size else 0.0
    atol = (rtol if scale_rtol is None else scale_rtol) * scale
    # NaN / Inf safe: "not within tolerance" (a NaN never satisfies <=), and a non-finite value must sit exactly
    # where the expectation has the same non-finite value
    with np.errstate(invalid="ignore"):
        err = np.abs(a - e)
        bad = ~(err <= rtol * np.abs(e) + atol)
    bad &= ~((a == e) & ~np.isnan(a))  # identical infinities are fine
    if bad.any():
        excess = np.where(np.isfinite(err), err - rtol * np.abs(e) - atol, np.inf)
        i = np.unravel_index(np.argmax(np.where(bad, excess, -np.inf)), a.shape)
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{a.size} elements out of tolerance; worst at {i}: "
            f"actual {a[i]!r} expected {e[i]!r} (|d|={abs(a[i]-e[i]):.3e}, scale {scale:.3e})")


def random_cameras(N, g, dt=None):
    """Seeded cameras for the transform tests: small rotations, ~425 px focal length with a little skew."""
    import torch as th
    dt = dt or th.float64
    ang = th.rand((N, 3), generator=g, dtype=dt) * 0.4 - 0.2
    cx, sx, cy, sy, cz, sz = ang[:, 0].cos(), ang[:, 0].sin(), ang[:, 1].cos(), ang[:, 1].sin(), ang[:, 2].cos(), ang[:, 2].sin()
    o, z = th.ones(N, dtype=dt), th.zeros(N, dtype=dt)
    Rx = th.stack((o, z, z, z, cx, -sx, z, sx, cx), -1).view(N, 3, 3)
    Ry = th.stack((cy, z, sy, z, o, z, -sy, z, cy), -1).view(N, 3, 3)
    Rz = th.stack((cz, -sz, z, sz, cz, z, z, z, o), -1).view(N, 3, 3)
    camrot = Rz @ Ry @ Rx
    campos = th.rand((N, 3), generator=g, dtype=dt) * 0.2 - 0.1
    focal = th.zeros((N, 2, 2), dtype=dt)
    focal[:, 0, 0] = 400 + 50 * th.rand(N, generator=g, dtype=dt)
    focal[:, 1, 1] = 400 + 50 * th.rand(N, generator=g, dtype=dt)
    focal[:, 0, 1] = 2 * th.rand(N, generator=g, dtype=dt)
    princpt = 256 + 10 * th.rand((N, 2), generator=g, dtype=dt)
    return campos, camrot, focal, princpt
