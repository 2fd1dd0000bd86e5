"""drtk.mipmap_grid_sample: grid_sample over a mip pyramid with trilinear + anisotropic filtering.

API mirror of `drtk/mipmap_grid_sample.py:18-127` (op `mipmap_grid_sampler_ext::mipmap_grid_sampler_2d`,
autograd in `src/mipmap_grid_sampler/mipmap_grid_sampler_module.cpp:44-196`): gradients go to every pyramid level
and to `grid`, never to `vt_dxdy_img`.  CUDA kernels: `csrc/samplers.cu` behind `drtk_b200_mipmap_grid_sample_*`.
"""
import ctypes
from typing import List, Optional

import torch as th
import torch.nn.functional as thf

from . import _lib
from . import _ops
from ._ops import _chk

_MODES = {"bilinear": 0, "bicubic": 2}
_PADS = {"zeros": 0, "border": 1, "reflection": 2}
_MAX_LEVELS = 11


def _level_arrays(levels):
    L = len(levels)
    ptrs = (ctypes.c_void_p * L)(*[t.data_ptr() for t in levels])
    hw = (ctypes.c_int64 * (2 * L))(*[d for t in levels for d in t.shape[2:]])
    st = (ctypes.c_int64 * (4 * L))(*[s for t in levels for s in t.stride()])
    return ptrs, hw, st


def _check(input, grid, vt_dxdy_img):
    who = "mipmap_aniso_grid_sampler_2d()"
    _chk(len(input) >= 1, f"{who}: expected input to have at least one mipmap level")
    _chk(len(input) <= _MAX_LEVELS, f"{who}: at most {_MAX_LEVELS} mipmap levels are supported")
    x0 = input[0]
    _chk(x0.device == grid.device and x0.is_cuda,
         f"{who}: expected input and grid to be on same device, but input is on {x0.device} and grid is on {grid.device}")
    _chk(x0.dtype == grid.dtype, f"{who}: expected input and grid to have same dtype, but input has {x0.dtype} and grid has {grid.dtype}")
    _chk(x0.dim() == 4 and grid.dim() == 4 and vt_dxdy_img.dim() == 5,
         f"{who}: expected 4D input and grid with same number of dimensions and 5D vt_dxdy_img, but got input with sizes "
         f"{tuple(x0.shape)} and grid with sizes {tuple(grid.shape)} and vt_dxdy_img with sizes {tuple(vt_dxdy_img.shape)}")
    _chk(x0.size(0) == grid.size(0) == vt_dxdy_img.size(0),
         f"{who}: expected grid, vt_dxdy_img and input to have same batch size, but got input with sizes {tuple(x0.shape)} "
         f"and grid with sizes {tuple(grid.shape)} and vt_dxdy_img with sizes {tuple(vt_dxdy_img.shape)}")
    _chk(grid.size(-1) == 2, f"{who}: expected grid to have size 2 in last dimension, but got grid with sizes {tuple(grid.shape)}")
    _chk(vt_dxdy_img.shape[-2:] == (2, 2) and vt_dxdy_img.shape[1:3] == grid.shape[1:3],
         f"{who}: expected vt_dxdy_img to have size 2 in last two dimension, but got grid with sizes {tuple(grid.shape)}")
    for t in input[1:]:
        _chk(t.device == x0.device and t.dtype == x0.dtype and t.dim() == 4 and t.shape[:2] == x0.shape[:2],
             f"{who}: expected all inputs to have same device, dtype, layout, and first two dimensions")
    _chk(all(d > 0 for t in input for d in t.shape[2:]),
         f"grid_sampler(): expected input to have non-empty spatial dimensions, but input has sizes {tuple(x0.shape)}")
    _chk(x0.dtype == th.float32, f"{who}: drtk_b200 computes in float32 only, but input has {x0.dtype}; cast it to float32")


class _MipmapGridSample(th.autograd.Function):
    @staticmethod
    def forward(ctx, grid, vt_dxdy_img, opts, *levels):
        max_aniso, pad, interp, align, force, clip = opts
        lib = _lib.load()
        N, C = levels[0].shape[:2]
        H, W = grid.shape[1:3]
        ptrs, hw, st = _level_arrays(levels)
        with th.cuda.device(grid.device):
            out = th.empty((N, C, H, W), dtype=th.float32, device=grid.device)
            rc = lib.drtk_b200_mipmap_grid_sample_forward(
                ptrs, hw, st, len(levels), _lib.ptr(grid), _lib.strides(grid), _lib.ptr(vt_dxdy_img),
                _lib.strides(vt_dxdy_img), N, C, H, W, max_aniso, pad, interp, int(align), int(force), int(clip),
                _lib.ptr(out), th.cuda.current_stream(grid.device).cuda_stream)
        _lib.check(rc, "mipmap_grid_sample()")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(grid, vt_dxdy_img, *levels)
        ctx.opts = opts
        return out

    @staticmethod
    def backward(ctx, grad_out):
        nlev = len(ctx.saved_tensors) - 2
        none = (None,) * (3 + nlev)
        need_tex = any(ctx.needs_input_grad[3:])
        if grad_out is None or not (need_tex or ctx.needs_input_grad[0]):
            return none
        grid, vt_dxdy_img, *levels = ctx.saved_tensors
        max_aniso, pad, interp, align, force, clip = ctx.opts
        lib = _lib.load()
        N, C = levels[0].shape[:2]
        H, W = grid.shape[1:3]
        grad_out = grad_out.float()
        ptrs, hw, st = _level_arrays(levels)
        with th.cuda.device(grid.device):
            # like the reference, the two gradients are produced together (one pass over the taps)
            g_levels = [th.empty(t.shape, dtype=th.float32, device=t.device) for t in levels]
            g_grid = th.empty(grid.shape, dtype=th.float32, device=grid.device)
            gptrs = (ctypes.c_void_p * nlev)(*[t.data_ptr() for t in g_levels])
            rc = lib.drtk_b200_mipmap_grid_sample_backward(
                _lib.ptr(grad_out), _lib.strides(grad_out), ptrs, hw, st, nlev, _lib.ptr(grid), _lib.strides(grid),
                _lib.ptr(vt_dxdy_img), _lib.strides(vt_dxdy_img), N, C, H, W, max_aniso, pad, interp, int(align),
                int(force), int(clip), gptrs, _lib.ptr(g_grid), th.cuda.current_stream(grid.device).cuda_stream)
        _lib.check(rc, "mipmap_grid_sample() backward")
        return (g_grid, None, None, *g_levels)


@th.compiler.disable
def mipmap_grid_sample(
    input: List[th.Tensor],
    grid: th.Tensor,
    vt_dxdy_img: th.Tensor,
    max_aniso: int,
    mode: str = "bilinear",
    padding_mode: str = "zeros",
    align_corners: Optional[bool] = None,
    force_max_aniso: Optional[bool] = False,
    clip_grad: Optional[bool] = False,
) -> th.Tensor:
    """Sample the pyramid `input` ([N,C,H_l,W_l], finest first; levels may be missing at the coarse end) at `grid`
    [N,H,W,2] in [-1,1], choosing levels and the number (<= max_aniso) and direction of the anisotropic samples
    from the uv Jacobian `vt_dxdy_img` [N,H,W,2,2] (uv in 0..1 units per pixel).  -> [N,C,H,W].

    mode 'bilinear' | 'bicubic'; padding_mode 'zeros' | 'border' | 'reflection'; `force_max_aniso` always takes
    max_aniso samples; `clip_grad` shrinks the footprint instead of spreading taps when the needed level is missing.
    See `drtk/mipmap_grid_sample.py:18-127`."""
    if mode not in _MODES:
        raise ValueError(f"mipmap_grid_sample(): only 'bilinear' and 'bicubic' modes are supported but got: '{mode}'")
    if padding_mode not in _PADS:
        raise ValueError("mipmap_grid_sample(): expected padding_mode to be 'zeros', 'border', or 'reflection', "
                         f"but got: '{padding_mode}'")
    input = list(input)
    if th.is_autocast_enabled():  # the reference's Autocast kernel casts everything to float32
        input = [_ops._autocast_one(t) for t in input]
        grid, vt_dxdy_img = _ops._autocast_one(grid), _ops._autocast_one(vt_dxdy_img)
    _check(input, grid, vt_dxdy_img)
    opts = (int(max_aniso), _PADS[padding_mode], _MODES[mode], bool(align_corners), bool(force_max_aniso), bool(clip_grad))
    _chk(opts[0] >= 1, "mipmap_grid_sample(): max_aniso must be at least 1")
    return _MipmapGridSample.apply(grid, vt_dxdy_img.float(), opts, *input)


def mipmap_grid_sample_ref(
    input: List[th.Tensor],
    grid: th.Tensor,
    vt_dxdy_img: th.Tensor,
    max_aniso: int,
    mode: str = "bilinear",
    padding_mode: str = "border",
    align_corners: Optional[bool] = False,
    high_quality: bool = False,
) -> th.Tensor:
    """Stock-torch statement of the op (any device / dtype), the counterpart of the reference's
    `mipmap_grid_sample_ref` (`drtk/mipmap_grid_sample.py:130-236`): agrees with `mipmap_grid_sample(...,
    force_max_aniso=True, clip_grad=False)` when `high_quality=False`.  `high_quality=True` takes the major axis
    of the footprint from an SVD of the Jacobian instead of the larger of its two rows."""
    levels = len(input)
    # uv derivatives -> texels: u scales with the width, v with the height, as in the kernel
    # (mipmap_grid_sampler_kernel.cu:452-453; the reference's own statement multiplies u by H and v by W,
    # drtk/mipmap_grid_sample.py:156-161, which only agrees with its kernel for square textures)
    size = th.as_tensor(input[0].shape[:1:-1], dtype=vt_dxdy_img.dtype, device=vt_dxdy_img.device)
    with th.no_grad():
        jac_px = vt_dxdy_img * size
        px, py = jac_px[..., 0, :].norm(dim=-1), jac_px[..., 1, :].norm(dim=-1)
        if high_quality:
            _, sv, vh = th.linalg.svd(jac_px)
            p_max, p_min = sv[..., 0], sv[..., 1]
            step = vh[..., 0, :] * sv[..., 0:1] / size
        else:
            p_max, p_min = th.max(px, py), th.min(px, py)
            step = th.where((px > py)[..., None], vt_dxdy_img[..., 0, :], vt_dxdy_img[..., 1, :])
        if max_aniso != 1:
            n = (p_max / p_min).ceil().clamp(max=max_aniso)
            n[n.isnan()] = 1
            lam = (p_max / n).log2()
        else:
            lam = p_max.log2()
        lam[lam.isinf()] = 0
        lam = lam.clamp(min=0, max=levels - 1 - 1e-6)
        d1 = lam.floor().long()
        frac = lam - d1.to(lam.dtype)
    sampled = []
    for tex in input:
        acc = 0
        for j in range(max_aniso):
            uv = grid + step * ((j + 1) / (max_aniso + 1) * 2.0 - 1.0) if max_aniso != 1 else grid
            acc = acc + thf.grid_sample(tex, uv, mode=mode, padding_mode=padding_mode, align_corners=align_corners)
        sampled.append(acc / max_aniso)
    if levels == 1:
        return sampled[0]
    stack = th.stack(sampled, 0)
    idx = th.stack((d1, d1 + 1), 0)[:, :, None].expand(-1, -1, stack.shape[2], -1, -1)
    lo, hi = th.gather(stack, 0, idx)
    return th.lerp(lo, hi, frac[:, None])
