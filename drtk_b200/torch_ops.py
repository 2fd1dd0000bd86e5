"""The dispatcher boundary: `torch.ops.drtk_b200_{rasterize,render,interpolate,edge_grad}_ext.*`.

`csrc/torch_shim.cpp` registers the reference's op schemas (`src/rasterize/rasterize_module.cpp:77-79`,
`src/render/render_module.cpp:89-91`, `src/interpolate/interpolate_module.cpp:632-634`,
`src/edge_grad/edge_grad_module.cpp:205-208`) under `drtk_b200_*_ext` with Autograd / Autocast / CUDA
implementations that call the C ABI of libdrtk_b200.so.  The shared object is built in-tree
(`drtk_b200/_torch_ops.so`, host C++ only) by `build()` and loaded with `torch.ops.load_library`, exactly how the
reference's `drtk/utils/load_torch_ops.py:14-28` loads its extensions.

Which host path the public functions (`drtk_b200.rasterize`, ...) take is decided once per process:
  DRTK_B200_DISPATCH=torch   -> these dispatcher ops (C++ autograd functions, no ctypes marshalling)
  DRTK_B200_DISPATCH=ctypes  -> the Python autograd functions over ctypes (`_ops.py`)
  unset                      -> torch when `_torch_ops.so` is present (it is after `build()`), else ctypes
`set_mode()` switches at run time (tools, tests).  Host cost of one forward + backward step on B200 (two-triangle 512^2
scene, the size of the reference's tutorials): 473 us over ctypes, 166 us over the dispatcher ops, the kernels the same.
Both end in the same C-ABI entry points and kernels.
"""
import os
import subprocess
import threading

import torch

from . import _lib

SO_PATH = os.path.join(_lib._HERE, "_torch_ops.so")
SRC = os.path.join(_lib.CSRC, "torch_shim.cpp")
_loaded = False
_lock = threading.Lock()


def available() -> bool:
    return os.path.exists(SO_PATH)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/torch_shim.cpp against this interpreter's torch (g++, no nvcc: the shim holds no device code) and
    link it to libdrtk_b200.so next to it ($ORIGIN rpath)."""
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    if (not force and os.path.exists(SO_PATH) and os.path.getmtime(SO_PATH) >= os.path.getmtime(SRC)
            and os.path.getmtime(SO_PATH) >= os.path.getmtime(os.path.join(_lib._HERE, "..", "include", "drtk_b200.h"))):
        return SO_PATH
    from torch.utils import cpp_extension as ce
    try:
        inc = ce.include_paths(device_type="cuda")
    except TypeError:  # older signature
        inc = ce.include_paths(cuda=True)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = ([gxx, "-O2", "-std=c++17", "-fPIC", "-shared", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
            SRC, "-o", SO_PATH] + [f"-I{i}" for i in inc] +
           [f"-L{libdir}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda", f"-L{_lib._HERE}", "-ldrtk_b200",
            "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{libdir}"])
    r = subprocess.run(cmd, capture_output=not verbose, text=True)
    if r.returncode != 0:
        raise RuntimeError("drtk_b200: building _torch_ops.so failed:\n" + (r.stderr or "")[-4000:])
    return SO_PATH


def load() -> None:
    """Register the ops with this process's dispatcher (idempotent).  Raises when the shared object is missing."""
    global _loaded
    if _loaded:
        return
    with _lock:
        if _loaded:
            return
        if not available():
            raise RuntimeError(f"drtk_b200: {SO_PATH} is missing; build it with `drtk_b200.torch_ops.build()`")
        _lib.load()  # libdrtk_b200.so first: fails loudly when the kernels are missing
        torch.ops.load_library(SO_PATH)
        _loaded = True


_mode = None


def enabled() -> bool:
    """True when the public API routes through the dispatcher ops (decided once per process)."""
    global _mode
    if _mode is None:
        want = os.environ.get("DRTK_B200_DISPATCH", "").strip().lower()
        if want not in ("", "torch", "ctypes"):
            raise RuntimeError(f"DRTK_B200_DISPATCH={want!r}: expected 'torch' or 'ctypes'")
        if want == "torch" or (want == "" and available()):
            load()
            _mode = "torch"
        else:
            _mode = "ctypes"
    return _mode == "torch"


def set_mode(mode: str) -> None:
    """Select the host path of the public functions for the rest of the process: "torch" or "ctypes"."""
    global _mode
    if mode not in ("torch", "ctypes"):
        raise ValueError(mode)
    if mode == "torch":
        load()
    _mode = mode


def rasterize(v, vi, height, width, wireframe=False):
    """-> [depth_img, index_img] (`drtk_b200_rasterize_ext::rasterize`)"""
    load()
    return torch.ops.drtk_b200_rasterize_ext.rasterize(v, vi, int(height), int(width), bool(wireframe))


def render(v, vi, index_img):
    """-> [depth_img, bary_img] (`drtk_b200_render_ext::render`)"""
    load()
    return torch.ops.drtk_b200_render_ext.render(v, vi, index_img)


def interpolate(vert_attributes, vi, index_img, bary_img):
    load()
    return torch.ops.drtk_b200_interpolate_ext.interpolate(vert_attributes, vi, index_img, bary_img)


def edge_grad_estimator(v_pix, v_pix_img, vi, img, index_img, max_dp_dr=1e4):
    load()
    return torch.ops.drtk_b200_edge_grad_ext.edge_grad_estimator(v_pix, v_pix_img, vi, img, index_img, float(max_dp_dr))


def edge_grad_estimator_fused(v_pix, vi, bary_img, img, index_img, max_dp_dr=1e4):
    load()
    return torch.ops.drtk_b200_edge_grad_ext.edge_grad_estimator_fused(v_pix, vi, bary_img, img, index_img, float(max_dp_dr))
