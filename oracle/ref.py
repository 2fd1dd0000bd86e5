"""Access to the UNMODIFIED reference kernels built by oracle/build_ref.py into oracle/_ref/.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline / --impl reference).
The shared objects hold the reference's TORCH_LIBRARY ops with both its CUDA kernels and its
CPU twins; which one runs is decided by the device of the input tensors, exactly as in the
reference (`src/*/..._module.cpp`, e.g. `src/render/render_module.cpp:43`).

The thin wrappers below repeat what the reference's Python layer does around the ops
(`drtk/rasterize.py:61-65`, `drtk/render.py:35-39`, `drtk/interpolate.py:47-50`,
`drtk/edge_grad_estimator.py:165-180`) because /root/reference does not exist on the GPU box.
"""
import os

import torch as th

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_NAMES = ("rasterize", "render", "interpolate", "edge_grad")
_SAMPLERS = ("grid_scatter", "mipmap_grid_sampler")
_loaded = False
_samplers_loaded = False


def available() -> bool:
    return all(os.path.exists(os.path.join(_REF, f"{n}_ext.so")) for n in _NAMES)


def load() -> None:
    global _loaded
    if _loaded:
        return
    if not available():
        raise RuntimeError("oracle/_ref/*.so missing: run `python oracle/build_ref.py` where /root/reference exists")
    for n in _NAMES:
        th.ops.load_library(os.path.join(_REF, f"{n}_ext.so"))
    _loaded = True


def _exp(vi, n):
    return vi[None].expand(n, -1, -1) if vi.ndim == 2 else vi


def rasterize_with_depth(v, vi, height, width, wireframe=False):
    load()
    depth_img, index_img = th.ops.rasterize_ext.rasterize(v, _exp(vi, v.shape[0]), height, width, wireframe)
    return depth_img, index_img


def rasterize(v, vi, height, width, wireframe=False):
    return rasterize_with_depth(v, vi, height, width, wireframe)[1]


def render(v, vi, index_img):
    load()
    depth_img, bary_img = th.ops.render_ext.render(v, _exp(vi, v.shape[0]), index_img)
    return depth_img, bary_img


def interpolate(vert_attributes, vi, index_img, bary_img):
    load()
    return th.ops.interpolate_ext.interpolate(vert_attributes, _exp(vi, vert_attributes.shape[0]), index_img, bary_img)


def edge_grad_estimator(v_pix, vi, bary_img, img, index_img, v_pix_img_hook=None, max_dp_dr=1e4):
    load()
    vi = _exp(vi, v_pix.shape[0])
    v_pix_img = interpolate(v_pix, vi, index_img, bary_img.detach())
    out = th.ops.edge_grad_ext.edge_grad_estimator(v_pix, v_pix_img, vi, img, index_img, max_dp_dr)
    if v_pix_img_hook is not None:
        v_pix_img.register_hook(v_pix_img_hook)
    return out


# ---- samplers: CUDA-only ops of the reference --------------------------------------
_MODE = {"bilinear": 0, "bicubic": 2}
_PAD = {"zeros": 0, "border": 1, "reflection": 2}


def samplers_available() -> bool:
    return all(os.path.exists(os.path.join(_REF, f"{n}_ext.so")) for n in _SAMPLERS)


def _load_samplers() -> None:
    global _samplers_loaded
    if not _samplers_loaded:
        if not samplers_available():
            raise RuntimeError("oracle/_ref sampler extensions missing: run `python oracle/build_ref.py`")
        for n in _SAMPLERS:
            th.ops.load_library(os.path.join(_REF, f"{n}_ext.so"))
        _samplers_loaded = True


def grid_scatter(input, grid, output_height, output_width, mode="bilinear", padding_mode="border", align_corners=None):
    """`drtk/grid_scatter.py:18-105` around `grid_scatter_ext::grid_scatter_2d`."""
    _load_samplers()
    return th.ops.grid_scatter_ext.grid_scatter_2d(input, grid, output_height, output_width, _PAD[padding_mode],
                                                   _MODE[mode], bool(align_corners))


def mipmap_grid_sample(input, grid, vt_dxdy_img, max_aniso, mode="bilinear", padding_mode="zeros", align_corners=None,
                       force_max_aniso=False, clip_grad=False):
    """`drtk/mipmap_grid_sample.py:18-127` around `mipmap_grid_sampler_ext::mipmap_grid_sampler_2d`."""
    _load_samplers()
    return th.ops.mipmap_grid_sampler_ext.mipmap_grid_sampler_2d(
        list(input), grid, vt_dxdy_img, max_aniso, _PAD[padding_mode], _MODE[mode], bool(align_corners),
        bool(force_max_aniso), bool(clip_grad))
