"""drtk.edge_grad_estimator on the B200 kernels.

API mirror of the reference `drtk/edge_grad_estimator.py:19-180` and of the autograd structure
of `EdgeGradEstimatorFunction` (`src/edge_grad/edge_grad_module.cpp:116-170`):

    v_pix --(conduit)--> v_pix_img [N,3,H,W] --+
                                               +--> edge_grad_estimator op --> img (unchanged)
    img ---------------------------------------+

* forward is the identity on `img`;
* backward turns dL/dimg at visibility discontinuities into dL/d(v_pix_img) (the B200 gather
  kernel), hands `grad_output` through to `img`, and dL/d(v_pix_img) flows on to `v_pix` through
  the backward of `interpolate` with C = 3 (bary detached);
* `v_pix_img_hook` is registered on `v_pix_img` and therefore observes (or replaces) the
  [N,3,H,W] gradient image, exactly like the reference.

Two launch plans produce that gradient:
* with a `v_pix_img_hook` (the image must exist): edge_grad kernel -> [N,3,H,W] -> C = 3
  interpolate-backward kernel, as in the reference;
* without a hook: ONE fused kernel (`drtk_b200_edge_grad_backward_fused`) that multiplies each non-zero
  pixel gradient by the pixel's barycentrics and REDs it into grad_v_pix directly.  Same sums, no image.

Difference to the reference (forward only, values identical): the reference materialises
`v_pix_img = interpolate(v_pix, ...)` although its values are never read
(`drtk/edge_grad_estimator.py:168-172`, 28 B/px of traffic).  Here the conduit's forward returns a
stride-0 placeholder of the right shape -- no kernel, no memory -- and only its backward (the
C = 3 interpolate-backward kernel) does work.
"""
from typing import Callable, Optional

import torch as th

from . import _ops, torch_ops


class _VPixImgConduit(th.autograd.Function):
    """interpolate(v_pix, vi, index_img, bary.detach()) as an autograd edge without the forward."""

    @staticmethod
    def forward(ctx, v_pix, vi, index_img, bary_img):
        ctx.save_for_backward(v_pix, vi, index_img, bary_img)
        ctx.set_materialize_grads(False)
        N, _, H, W = bary_img.shape
        return th.empty((), dtype=v_pix.dtype, device=v_pix.device).expand(N, 3, H, W)

    @staticmethod
    def backward(ctx, grad_v_pix_img):
        if grad_v_pix_img is None or not ctx.needs_input_grad[0]:
            return None, None, None, None
        v_pix, vi, index_img, bary_img = ctx.saved_tensors
        native = (th.float32, th.float64)
        v32 = v_pix.detach() if v_pix.dtype in native else v_pix.detach().float()
        b32 = bary_img.detach().to(v32.dtype)
        ga, _ = _ops.interpolate_backward(grad_v_pix_img, v32, vi, index_img, b32, True, False)
        return ga.to(v_pix.dtype), None, None, None


class _EdgeGradFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, v_pix, v_pix_img, vi, img, index_img, max_dp_dr):
        _ops.check_edge_grad(v_pix, v_pix_img, vi, img, index_img)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(v_pix, img, index_img, vi)
        ctx.max_dp_dr = float(max_dp_dr)
        return img.view_as(img)

    @staticmethod
    def backward(ctx, grad_output):
        if grad_output is None:
            return None, None, None, None, None, None
        if not ctx.needs_input_grad[1]:  # v_pix_img does not require grad (:141-149)
            return None, None, None, grad_output, None, None
        v_pix, img, index_img, vi = ctx.saved_tensors
        grad_v_pix_img = _ops.edge_grad_backward(v_pix.detach(), img.detach(), index_img, vi, grad_output,
                                                 ctx.max_dp_dr)
        return None, grad_v_pix_img, None, grad_output, None, None


class _EdgeGradFusedFn(th.autograd.Function):
    """edge_grad_estimator op + conduit in one node (no hook => nobody can observe v_pix_img)."""

    @staticmethod
    def forward(ctx, v_pix, vi, bary_img, img, index_img, max_dp_dr):
        _ops.check_edge_grad(v_pix, bary_img, vi, img, index_img)
        _ops._check_interp(v_pix, vi, index_img, bary_img)
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(v_pix, img, index_img, vi, bary_img)
        ctx.max_dp_dr = float(max_dp_dr)
        return img.view_as(img)

    @staticmethod
    def backward(ctx, grad_output):
        if grad_output is None:
            return None, None, None, None, None, None
        if not ctx.needs_input_grad[0]:
            return None, None, None, grad_output, None, None
        v_pix, img, index_img, vi, bary_img = ctx.saved_tensors
        gv = _ops.edge_grad_backward_fused(v_pix.detach(), img.detach(), index_img, vi, grad_output,
                                           bary_img, ctx.max_dp_dr)
        return gv.to(v_pix.dtype), None, None, grad_output, None, None


@th.compiler.disable
def edge_grad_estimator(
    v_pix: th.Tensor,
    vi: th.Tensor,
    bary_img: th.Tensor,
    img: th.Tensor,
    index_img: th.Tensor,
    v_pix_img_hook: Optional[Callable[[th.Tensor], None]] = None,
    max_dp_dr: float = 1e4,
) -> th.Tensor:
    """Make `img` differentiable w.r.t. `v_pix` at visibility discontinuities.

    Args (as in the reference): v_pix [N,V,3] pixel-space vertices (camera-space z); vi [F,3] or
    [N,F,3] int32; bary_img [N,3,H,W]; img [N,C,H,W]; index_img [N,H,W] int32; optional backward
    hook on the [N,3,H,W] image-space gradient; max_dp_dr clamp for intersecting geometry
    (<= 0 disables the clamp).  Returns `img` unchanged but requiring grad.
    """
    if vi.ndim == 2:
        vi = vi[None, ...].expand(v_pix.shape[0], -1, -1)
    if torch_ops.enabled():  # dispatcher ops (C++ autograd functions); the zero-cost conduit stays a Python node
        if v_pix_img_hook is None:
            return torch_ops.edge_grad_estimator_fused(v_pix, vi, bary_img.detach(), img, index_img, max_dp_dr)
        v_pix_c, bary_c = _ops.autocast_f32(v_pix, bary_img)
        v_pix_img = _VPixImgConduit.apply(v_pix_c, vi, index_img, bary_c.detach())
        out = torch_ops.edge_grad_estimator(v_pix, v_pix_img, vi, img, index_img, max_dp_dr)
        if v_pix_img.requires_grad:
            v_pix_img.register_hook(v_pix_img_hook)
        return out
    v_pix, bary_img, img = _ops.autocast_f32(v_pix, bary_img, img)  # (src/edge_grad/edge_grad_module.cpp:172-196)
    if v_pix_img_hook is None:
        return _EdgeGradFusedFn.apply(v_pix, vi, bary_img.detach(), img, index_img, max_dp_dr)
    v_pix_img = _VPixImgConduit.apply(v_pix, vi, index_img, bary_img.detach())
    out = _EdgeGradFn.apply(v_pix, v_pix_img, vi, img, index_img, max_dp_dr)
    if v_pix_img_hook is not None and v_pix_img.requires_grad:
        v_pix_img.register_hook(v_pix_img_hook)
    return out
