"""Loader and ctypes prototypes for libdrtk_b200.so (the C ABI declared in include/drtk_b200.h).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is
raised.  `build()` compiles the library in-tree with nvcc for sm_100a (no GPU required).
"""
import ctypes
import os
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# DRTK_B200_LIB: developer override for A/B runs of kernel variants (tools/build_variants.sh)
LIB_PATH = os.environ.get("DRTK_B200_LIB") or os.path.join(_HERE, "libdrtk_b200.so")
CSRC = os.path.join(_HERE, "csrc")

ABI_VERSION = 2  # DRTK_B200_ABI_VERSION of include/drtk_b200.h

_lib = None
_lock = threading.Lock()

_P = ctypes.c_void_p
_I64 = ctypes.c_int64
_INT = ctypes.c_int
_F32 = ctypes.c_float
_F64 = ctypes.c_double
_SZ = ctypes.c_size_t

# name -> (restype, argtypes); kept in sync with include/drtk_b200.h (tests/test_abi.py checks
# that every prototype declared in the header is exported and listed here).
PROTOTYPES = {
    "drtk_b200_abi_version": (_INT, []),
    "drtk_b200_error_string": (ctypes.c_char_p, [_INT]),
    "drtk_b200_rasterize_workspace_bytes": (_SZ, [_I64, _I64, _I64, _I64, _INT]),
    "drtk_b200_rasterize": (
        _INT,
        [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _P, _P, _P, _SZ, _P],
    ),
    "drtk_b200_render_forward": (
        _INT,
        [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P],
    ),
    "drtk_b200_render_backward_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "drtk_b200_render_backward": (
        _INT,
        [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _SZ, _P],
    ),
    "drtk_b200_interpolate_forward": (
        _INT,
        [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _P],
    ),
    "drtk_b200_interpolate_backward_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "drtk_b200_interpolate_backward": (
        _INT,
        [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _P, _P, _SZ, _P],
    ),
    "drtk_b200_edge_grad_backward": (
        _INT,
        [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _F32, _P, _P],
    ),
    "drtk_b200_edge_grad_backward_fused_workspace_bytes": (_SZ, [_I64, _I64]),
    "drtk_b200_interpolation_matrix": (_INT, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "drtk_b200_interpolation_matrix_backward": (_INT, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_interpolation_normal_matrix_values": (_INT, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_interpolation_normal_matrix_values_backward": (_INT, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_mipmap_grid_sample_forward": (
        _INT, [_P, _P, _P, _INT, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _INT, _INT, _P, _P]),
    "drtk_b200_mipmap_grid_sample_backward": (
        _INT, [_P, _P, _P, _P, _P, _INT, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _INT, _INT, _INT,
               _P, _P, _P]),
    "drtk_b200_grid_scatter_forward": (
        _INT, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _P, _P]),
    "drtk_b200_grid_scatter_backward": (
        _INT, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _INT, _INT, _INT, _P, _P, _P]),
    "drtk_b200_rasterize_f64_workspace_bytes": (_SZ, [_I64, _I64, _I64]),
    "drtk_b200_rasterize_f64": (_INT, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _INT, _P, _P, _P, _SZ, _P]),
    "drtk_b200_render_forward_f64": (_INT, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "drtk_b200_render_backward_f64": (
        _INT, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_interpolate_forward_f64": (
        _INT, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_interpolate_backward_f64": (
        _INT, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P, _P, _P]),
    "drtk_b200_edge_grad_backward_f64": (
        _INT, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _F64, _P, _P]),
    "drtk_b200_screen_space_uv_derivative": (
        _INT, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_transform_forward": (_INT, [_P, _P, _P, _P, _INT, _INT, _I64, _I64, _P, _P, _P]),
    "drtk_b200_transform_backward": (
        _INT, [_P, _P, _P, _P, _INT, _INT, _P, _P, _P, _P, _I64, _I64, _P, _P, _P]),
    "drtk_b200_batch_sum": (_INT, [_P, _I64, _I64, _I64, _P, _P]),
    "drtk_b200_batch_sum_allreduce_grid": (_INT, []),
    "drtk_b200_batch_sum_allreduce": (_INT, [_P, _I64, _I64, _I64, _P, _P, _P, _INT, _INT, ctypes.c_uint32, _P, _INT, _P]),
    "drtk_b200_edge_grad_backward_fused": (
        _INT,
        [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _F32, _P, _P, _SZ, _P],
    ),
}


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libdrtk_b200.so in-tree (nvcc, sm_100a).  Cross-compiles without a GPU."""
    args = ["make", "-C", CSRC, "-j8"]
    if force:
        args.append("-B")
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(args, stdout=out)
    return LIB_PATH


def load():
    """Return the loaded CDLL; raises RuntimeError when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"drtk_b200: native library {LIB_PATH} is missing. Build it with "
                "`python -c 'import drtk_b200; drtk_b200.build()'` (nvcc, sm_100a). "
                "There is no CPU or PyTorch fallback."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.drtk_b200_abi_version() != ABI_VERSION:
            raise RuntimeError("drtk_b200: ABI version mismatch between python host and native library")
        _lib = lib
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = load().drtk_b200_error_string(code)
        raise RuntimeError(f"{what}: {msg.decode() if msg else code}")


def strides(t):
    """Host array of element strides, as the C ABI expects."""
    s = t.stride()
    return (ctypes.c_int64 * len(s))(*s)


def ptr(t):
    return None if t is None else t.data_ptr()
