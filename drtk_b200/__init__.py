"""drtk_b200 -- the DRTK rasterisation hot path (rasterize / render / interpolate /
edge_grad_estimator, forward + backward) on hand-written sm_100a CUDA kernels.

Drop-in for the same-named functions of facebookresearch/DRTK (`drtk/__init__.py:8-33`):

    import drtk_b200 as drtk            # or: drtk_b200.install_as_drtk(); import drtk
    index_img = drtk.rasterize(v_pix, vi, height=H, width=W)
    depth_img, bary_img = drtk.render(v_pix, vi, index_img)
    img = drtk.interpolate(attrs, vi, index_img, bary_img)
    img = drtk.edge_grad_estimator(v_pix, vi, bary_img, img, index_img)

The kernels live in libdrtk_b200.so behind the C ABI of include/drtk_b200.h; this package is
the Python host (argument checks, allocation, autograd).  There is no CPU or eager fallback:
without the native library every op raises.
"""
import sys

from . import _lib
from . import torch_ops  # noqa: F401
from ._lib import build  # noqa: F401
from .edge_grad_estimator import edge_grad_estimator  # noqa: F401
from .grid_scatter import grid_scatter, grid_scatter_ref  # noqa: F401
from .interpolate import interpolate, interpolate_ref, interpolation_matrix, interpolation_normal_matrix  # noqa: F401
from .mipmap_grid_sample import mipmap_grid_sample, mipmap_grid_sample_ref  # noqa: F401
from .rasterize import rasterize, rasterize_with_depth  # noqa: F401
from .render import render, render_ref  # noqa: F401
from .transform import transform, transform_with_v_cam  # noqa: F401
from . import utils  # noqa: F401,E402
from .screen_space_uv_derivative import screen_space_uv_derivative  # noqa: F401,E402

__version__ = "0.1.0"

__all__ = [
    "rasterize", "rasterize_with_depth", "render", "interpolate", "interpolation_matrix", "interpolation_normal_matrix",
    "edge_grad_estimator", "render_ref", "interpolate_ref", "grid_scatter", "grid_scatter_ref", "mipmap_grid_sample",
    "mipmap_grid_sample_ref", "screen_space_uv_derivative",
    "transform", "transform_with_v_cam", "utils", "build", "torch_ops", "install_as_drtk", "native_library_path",
]


def native_library_path() -> str:
    return _lib.LIB_PATH


def install_as_drtk() -> None:
    """Make `import drtk` (and `from drtk.render import render`, ...) resolve to this package,
    for pipelines written against the reference."""
    this = sys.modules[__name__]
    if "drtk" in sys.modules and sys.modules["drtk"] is not this:
        raise RuntimeError("a different `drtk` package is already imported")
    sys.modules["drtk"] = this
    for sub in ("rasterize", "render", "interpolate", "edge_grad_estimator", "transform", "utils", "grid_scatter",
                "mipmap_grid_sample", "screen_space_uv_derivative"):
        sys.modules[f"drtk.{sub}"] = sys.modules[f"{__name__}.{sub}"]
