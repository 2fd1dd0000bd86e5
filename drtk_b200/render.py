"""drtk.render on the B200 kernels (API mirror of the reference `drtk/render.py:17-39`).

Autograd contract of the reference's RenderFunction (`src/render/render_module.cpp:27-72`):
gradient flows to `v` only, and only if `v` required grad at forward time.
"""
from typing import Tuple

import torch as th

from . import _ops, torch_ops


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
        vd = v.detach() if v.dtype in (th.float32, th.float64) else v.detach().float()
        grad_v = _ops.render_backward(vd, vi, index_img, grad_depth, grad_bary)
        return grad_v.to(v.dtype), None, None


@th.compiler.disable
def render(v: th.Tensor, vi: th.Tensor, index_img: th.Tensor) -> Tuple[th.Tensor, th.Tensor]:
    """Per-pixel depth and perspective-correct barycentrics.

    Args: v [N,V,3]; vi [F,3] or [N,F,3] int32; index_img [N,H,W] int32 (from rasterize).
    Returns: depth_img [N,H,W], bary_img [N,3,H,W] (planar, like the reference kernel output).
    """
    if vi.ndim == 2:
        vi = vi[None].expand(v.shape[0], -1, -1)
    if torch_ops.enabled():
        depth_img, bary_img = torch_ops.render(v, vi, index_img)
        return depth_img, bary_img
    (v,) = _ops.autocast_f32(v)
    return _RenderFn.apply(v, vi, index_img)


def render_ref(v: th.Tensor, vi: th.Tensor, index_img: th.Tensor):
    """Pure-PyTorch, float64, differentiable statement of :func:`render` (any device), for tests and debugging -- the
    counterpart of the reference's `render_ref` (`drtk/render.py:66-132`).  `vi` is [F,3].
    Returns (depth_img [N,H,W], bary_img [N,3,H,W]); zeros where index_img == -1."""
    dt = v.dtype
    v64 = v.double()
    N, H, W = index_img.shape
    mask = index_img != -1
    tri = index_img.clamp(min=0).long()
    corners = vi.long()[tri]                                                          # [N,H,W,3]
    p = th.gather(v64, 1, corners.reshape(N, -1, 1).expand(-1, -1, 3)).reshape(N, H, W, 3, 3)   # [.., corner, xyz]
    p0, p1, p2 = p[..., 0, :], p[..., 1, :], p[..., 2, :]

    def epsclamp(x):
        return th.where(x < 0, x.clamp(max=-1e-16), x.clamp(min=1e-16))

    ys, xs = th.meshgrid(th.arange(H, device=v.device, dtype=th.float64), th.arange(W, device=v.device, dtype=th.float64),
                         indexing="ij")
    e01, e02 = p1 - p0, p2 - p0
    den = epsclamp(e01[..., 0] * e02[..., 1] - e01[..., 1] * e02[..., 0])
    qx, qy = xs[None] - p0[..., 0], ys[None] - p0[..., 1]
    l1 = (qx * e02[..., 1] - qy * e02[..., 0]) / den
    l2 = (qy * e01[..., 0] - qx * e01[..., 1]) / den
    l0 = 1.0 - l1 - l2
    w0, w1, w2 = l0 / epsclamp(p0[..., 2]), l1 / epsclamp(p1[..., 2]), l2 / epsclamp(p2[..., 2])
    depth = 1.0 / epsclamp(w0 + w1 + w2)
    bary = th.stack((w0 * depth, w1 * depth, w2 * depth), 1) * mask[:, None]
    return (depth * mask).to(dt), bary.to(dt)
