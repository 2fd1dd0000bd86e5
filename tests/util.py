"""Shared comparison helpers for the test-suite."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_cases():
    # wire_*.npz (wireframe, reference CUDA outputs) and mat_*.npz (interpolation matrices, reference CPU outputs)
    # are handled by their own tests
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN, "*.npz"))
                  if not os.path.basename(p).startswith(("wire_", "mat_")))


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
    bad = np.abs(a - e) > rtol * np.abs(e) + atol
    if bad.any():
        i = np.unravel_index(np.argmax(np.abs(a - e) - rtol * np.abs(e) - atol), a.shape)
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{a.size} elements out of tolerance; worst at {i}: "
            f"actual {a[i]!r} expected {e[i]!r} (|d|={abs(a[i]-e[i]):.3e}, scale {scale:.3e})")
