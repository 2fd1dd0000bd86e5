"""drtk.render on the B200 kernels (API mirror of the reference `drtk/render.py:17-39`).

Autograd contract of the reference's RenderFunction (`src/render/render_module.cpp:27-72`):
gradient flows to `v` only, and only if `v` required grad at forward time.
"""
from typing import Tuple

import torch as th

from . import _ops


class _RenderFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, v, vi, index_img):
        depth_img, bary_img = _ops.render_forward(v, vi, index_img)
        ctx.save_for_backward(v, vi, index_img)
        # the reference materialises undefined grads as zeros; skipping that is value-identical
        ctx.set_materialize_grads(False)
        return depth_img, bary_img

    @staticmethod
    def backward(ctx, grad_depth, grad_bary):
        if not ctx.needs_input_grad[0]:
            return None, None, None
        v, vi, index_img = ctx.saved_tensors
        grad_v = _ops.render_backward(v.detach().float() if v.dtype != th.float32 else v.detach(), vi,
                                      index_img, grad_depth, grad_bary)
        return grad_v.to(v.dtype), None, None


@th.compiler.disable
def render(v: th.Tensor, vi: th.Tensor, index_img: th.Tensor) -> Tuple[th.Tensor, th.Tensor]:
    """Per-pixel depth and perspective-correct barycentrics.

    Args: v [N,V,3]; vi [F,3] or [N,F,3] int32; index_img [N,H,W] int32 (from rasterize).
    Returns: depth_img [N,H,W], bary_img [N,3,H,W] (planar, like the reference kernel output).
    """
    if vi.ndim == 2:
        vi = vi[None].expand(v.shape[0], -1, -1)
    return _RenderFn.apply(v, vi, index_img)
