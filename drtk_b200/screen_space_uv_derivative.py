"""drtk.screen_space_uv_derivative: per-pixel Jacobian of the texture coordinates w.r.t. the pixel position, the
`vt_dxdy_img` argument of `mipmap_grid_sample`.

API mirror of `drtk/screen_space_uv_derivative.py:16-80`.  On CUDA float32 inputs without a gradient request the whole
chain runs as ONE kernel (`csrc/uv_derivative.cu`, C ABI `drtk_b200_screen_space_uv_derivative`); when a gradient is
needed (or for float64) the differentiable composition of the reference is used, built on this package's
`interpolate` and `project_points_grad`.
"""
from typing import Optional, Sequence

import torch as th

from . import _lib
from .interpolate import interpolate
from .utils import face_dpdt, project_points_grad


def _composed(v, vt, vi, vti, index_img, bary_img, mask, campos, camrot, focal, dist_mode, dist_coeff):
    N, (H, W) = v.shape[0], index_img.shape[1:]
    dpdt_t, vf = face_dpdt(v, vt, vi.long(), vti.long())                     # [N,F,2,3], [N,F,3,3]
    F = dpdt_t.shape[1]
    # a mesh of 3F unshared vertices: per-face constants must not be blended across faces
    vi_dis = th.arange(0, 3 * F, dtype=th.int32, device=v.device).view(-1, 3)
    per_corner = dpdt_t[:, :, None].expand(-1, -1, 3, -1, -1).reshape(N, 3 * F, 6)
    dpdt_img = interpolate(per_corner, vi_dis, index_img, bary_img).permute(0, 2, 3, 1).reshape(N, H, W, 2, 3)
    p_img = interpolate(vf.reshape(N, 3 * F, 3), vi_dis, index_img, bary_img).permute(0, 2, 3, 1)
    p_img = p_img[:, :, :, None].expand(-1, -1, -1, 2, -1)
    dpix = project_points_grad(dpdt_img.reshape(N, -1, 3), p_img.reshape(N, -1, 3), campos, camrot, focal, dist_mode,
                               dist_coeff).view(N, H, W, 2, 2)               # [..., i, j] = d pix_j / d t_i
    out, _ = th.linalg.inv_ex(dpix)
    return th.where(mask[..., None, None], out, th.zeros_like(out))


@th.compiler.disable
def screen_space_uv_derivative(
    v: th.Tensor,
    vt: th.Tensor,
    vi: th.Tensor,
    vti: th.Tensor,
    index_img: th.Tensor,
    bary_img: th.Tensor,
    mask: th.Tensor,
    campos: th.Tensor,
    camrot: th.Tensor,
    focal: th.Tensor,
    dist_mode: Optional[Sequence[str]] = None,
    dist_coeff: Optional[th.Tensor] = None,
) -> th.Tensor:
    """v [N,V,3] world-space vertices, vt [N,T,2] uv coordinates, vi / vti [F,3] face indices into v / vt,
    index_img [N,H,W], bary_img [N,3,H,W], mask [N,H,W] bool, pinhole camera (campos [N,3], camrot [N,3,3],
    focal [N,2,2]).  Returns [N,H,W,2,2] = [[du/dx, dv/dx], [du/dy, dv/dy]], zero outside `mask`.
    Pixels without a triangle yield zeros on the fused path (the composition evaluates `interpolate`'s background
    sweep there -- meaningless values the mask is expected to remove)."""
    if dist_mode is not None:  # as the reference: distorted cameras are not implemented (projection.py:693-696)
        return _composed(v, vt, vi, vti, index_img, bary_img, mask, campos, camrot, focal, dist_mode, dist_coeff)
    floats = (v, vt, bary_img, campos, camrot, focal)
    needs_grad = th.is_grad_enabled() and any(t.requires_grad for t in floats)
    if needs_grad or not v.is_cuda or any(t.dtype != th.float32 for t in floats) or index_img.dtype != th.int32:
        return _composed(v, vt, vi, vti, index_img, bary_img, mask, campos, camrot, focal, None, None)
    N, H, W = index_img.shape
    lib = _lib.load()
    vi32 = vi if vi.dtype == th.int32 else vi.int()
    vti32 = vti if vti.dtype == th.int32 else vti.int()
    with th.cuda.device(v.device):
        cam = th.cat((campos.reshape(N, 3), camrot.reshape(N, 9), focal.reshape(N, 4)), 1).contiguous()
        m8 = mask.view(th.uint8) if mask.dtype == th.bool else mask.ne(0).view(th.uint8)
        out = th.empty((N, H, W, 2, 2), dtype=th.float32, device=v.device)
        rc = lib.drtk_b200_screen_space_uv_derivative(
            _lib.ptr(v), _lib.strides(v), _lib.ptr(vt), _lib.strides(vt), _lib.ptr(vi32), _lib.strides(vi32),
            _lib.ptr(vti32), _lib.strides(vti32), _lib.ptr(index_img), _lib.strides(index_img), _lib.ptr(bary_img),
            _lib.strides(bary_img), _lib.ptr(m8), _lib.strides(m8), _lib.ptr(cam), N, H, W, _lib.ptr(out),
            th.cuda.current_stream(v.device).cuda_stream)
    _lib.check(rc, "screen_space_uv_derivative()")
    return out
