"""drtk.interpolate on the B200 kernels (API mirror of the reference `drtk/interpolate.py:19-50`).

Autograd contract of the reference's InterpolateFunction
(`src/interpolate/interpolate_module.cpp:378-425`): gradients to `vert_attributes` and/or
`bary_img` according to their requires_grad; an undefined upstream gradient yields no gradients.
"""
import torch as th

from . import _ops


class _InterpolateFn(th.autograd.Function):
    @staticmethod
    def forward(ctx, vert_attributes, vi, index_img, bary_img):
        out = _ops.interpolate_forward(vert_attributes, vi, index_img, bary_img)
        ctx.save_for_backward(vert_attributes, vi, index_img, bary_img)
        ctx.set_materialize_grads(False)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        need_attr, _, _, need_bary = ctx.needs_input_grad
        if grad_out is None or not (need_attr or need_bary):
            return None, None, None, None
        attr, vi, index_img, bary_img = ctx.saved_tensors
        a32 = attr.detach() if attr.dtype == th.float32 else attr.detach().float()
        b32 = bary_img.detach() if bary_img.dtype == th.float32 else bary_img.detach().float()
        ga, gb = _ops.interpolate_backward(grad_out, a32, vi, index_img, b32, need_attr, need_bary)
        if ga is not None:
            ga = ga.to(attr.dtype)
        if gb is not None:
            gb = gb.to(bary_img.dtype)
        return ga, None, None, gb


@th.compiler.disable
def interpolate(vert_attributes: th.Tensor, vi: th.Tensor, index_img: th.Tensor, bary_img: th.Tensor) -> th.Tensor:
    """Barycentric interpolation of vertex attributes.

    Args: vert_attributes [N,V,C]; vi [F,3] or [N,F,3] int32; index_img [N,H,W] int32;
    bary_img [N,3,H,W].  Returns [N,C,H,W].  Pixels with index -1 receive the reference's
    coordinate sweep (non-zero values that must be ignored), see
    `src/interpolate/interpolate_kernel.cu:104-109`.
    """
    if vi.ndim == 2:
        vi = vi[None].expand(vert_attributes.shape[0], -1, -1)
    return _InterpolateFn.apply(vert_attributes, vi, index_img, bary_img)
