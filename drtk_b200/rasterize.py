"""drtk.rasterize / drtk.rasterize_with_depth on the B200 kernels.

API mirror of the reference `drtk/rasterize.py:16-103`: same names, argument meaning, return
values and the [F,3] -> [N,F,3] broadcast of `vi` (a stride-0 expand, never materialised).
Not differentiable (reference: outputs marked non-differentiable,
`src/rasterize/rasterize_module.cpp:40-51`); gradients come from `edge_grad_estimator`.
"""
from typing import Tuple

import torch as th

from . import _ops, torch_ops


@th.compiler.disable
def rasterize(v: th.Tensor, vi: th.Tensor, height: int, width: int, wireframe: bool = False) -> th.Tensor:
    """Rasterize the mesh (v [N,V,3] pixel-space xy + camera-space z, vi [F,3] or [N,F,3] int32).

    Returns index_img int32 [N,H,W]: id of the nearest covering triangle per pixel, -1 where empty.
    Pixel centres are at integer coordinates; the canvas spans (-0.5,-0.5)..(W-0.5,H-0.5).
    Bit-exact with the reference CUDA kernels.
    """
    if vi.ndim == 2:
        vi = vi[None].expand(v.shape[0], -1, -1)
    if torch_ops.enabled():
        return torch_ops.rasterize(v, vi, height, width, wireframe)[1]
    with th.no_grad():
        (v,) = _ops.autocast_f32(v.detach())
        _, index_img = _ops.rasterize(v, vi, height, width, wireframe)
    return index_img


@th.compiler.disable
def rasterize_with_depth(
    v: th.Tensor, vi: th.Tensor, height: int, width: int, wireframe: bool = False
) -> Tuple[th.Tensor, th.Tensor]:
    """Same as :func:`rasterize` but returns (depth_img f32 [N,H,W], index_img); depth is 0 where
    empty and is not differentiable (use :func:`drtk_b200.render` for differentiable depth)."""
    if vi.ndim == 2:
        vi = vi[None].expand(v.shape[0], -1, -1)
    if torch_ops.enabled():
        depth_img, index_img = torch_ops.rasterize(v, vi, height, width, wireframe)
        return depth_img, index_img
    with th.no_grad():
        (v,) = _ops.autocast_f32(v.detach())
        depth_img, index_img = _ops.rasterize(v, vi, height, width, wireframe)
    return depth_img, index_img
