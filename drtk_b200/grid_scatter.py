"""drtk.grid_scatter: the splatting transpose of grid_sample.

API mirror of `drtk/grid_scatter.py:18-105` (op `grid_scatter_ext::grid_scatter_2d`, autograd in
`src/grid_scatter/grid_scatter_module.cpp:35-112`).  CUDA kernels: `csrc/samplers.cu` behind
`drtk_b200_grid_scatter_forward / _backward`.
"""
from typing import Optional

import torch as th
import torch.nn.functional as thf

from . import _lib
from . import _ops
from ._ops import _chk

_MODES = {"bilinear": 0, "bicubic": 2}
_PADS = {"zeros": 0, "border": 1, "reflection": 2}


class _GridScatter(th.autograd.Function):
    @staticmethod
    def forward(ctx, input, grid, out_h, out_w, pad, interp, align):
        lib = _lib.load()
        N, C, H, W = input.shape
        with th.cuda.device(input.device):
            out = th.empty((N, C, out_h, out_w), dtype=th.float32, device=input.device)
            rc = lib.drtk_b200_grid_scatter_forward(
                _lib.ptr(input), _lib.strides(input), _lib.ptr(grid), _lib.strides(grid), N, C, H, W, out_h, out_w,
                pad, interp, int(align), _lib.ptr(out), th.cuda.current_stream(input.device).cuda_stream)
        _lib.check(rc, "grid_scatter()")
        ctx.set_materialize_grads(False)
        ctx.save_for_backward(input, grid)
        ctx.opts = (out_h, out_w, pad, interp, align)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        need_in, need_grid = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if grad_out is None or not (need_in or need_grid):
            return (None,) * 7
        input, grid = ctx.saved_tensors
        out_h, out_w, pad, interp, align = ctx.opts
        lib = _lib.load()
        N, C, H, W = input.shape
        grad_out = grad_out.float()
        with th.cuda.device(input.device):
            g_in = th.empty((N, C, H, W), dtype=th.float32, device=input.device) if need_in else None
            g_grid = th.empty((N, H, W, 2), dtype=th.float32, device=input.device) if need_grid else None
            rc = lib.drtk_b200_grid_scatter_backward(
                _lib.ptr(grad_out), _lib.strides(grad_out), _lib.ptr(input), _lib.strides(input), _lib.ptr(grid),
                _lib.strides(grid), N, C, H, W, out_h, out_w, pad, interp, int(align), _lib.ptr(g_in),
                _lib.ptr(g_grid), th.cuda.current_stream(input.device).cuda_stream)
        _lib.check(rc, "grid_scatter() backward")
        return g_in, g_grid, None, None, None, None, None


@th.compiler.disable
def grid_scatter(
    input: th.Tensor,
    grid: th.Tensor,
    output_height: int,
    output_width: int,
    mode: str = "bilinear",
    padding_mode: str = "border",
    align_corners: Optional[bool] = None,
) -> th.Tensor:
    """Every input pixel `input[n,:,h,w]` is splatted to the location `grid[n,h,w]` ([-1,1], grid_sample
    conventions) of an `[N,C,output_height,output_width]` image with bilinear / bicubic weights; contributions
    accumulate.  The forward is the input-gradient of grid_sample; the backward samples.  See
    `drtk/grid_scatter.py:18-105`."""
    if mode not in _MODES:
        raise ValueError(f"grid_scatter(): only 'bilinear' and 'bicubic' modes are supported but got: '{mode}'")
    if padding_mode not in _PADS:
        raise ValueError("grid_scatter(): expected padding_mode to be 'zeros', 'border', or 'reflection', "
                         f"but got: '{padding_mode}'")
    if th.is_autocast_enabled():
        input, grid = _ops._autocast_one(input), _ops._autocast_one(grid)
    who = "grid_scatter_2d()"
    _chk(input.device == grid.device and input.is_cuda,
         f"{who}: expected input and grid to be on same device, but input is on {input.device} and grid is on {grid.device}")
    _chk(input.dtype == grid.dtype, f"{who}: expected input and grid to have same dtype, but input has {input.dtype} and grid has {grid.dtype}")
    _chk(input.dim() == 4 and grid.dim() == 4,
         f"{who}: expected 4D input and grid with same number of dimensions, but got input with sizes {tuple(input.shape)} and grid with sizes {tuple(grid.shape)}")
    _chk(input.size(0) == grid.size(0) and input.shape[2:] == grid.shape[1:3],
         f"{who}: expected grid and input to have same batch size and spatial size, but got input with sizes {tuple(input.shape)} and grid with sizes {tuple(grid.shape)}")
    _chk(grid.size(-1) == 2, f"{who}: expected grid to have size 2 in last dimension, but got grid with sizes {tuple(grid.shape)}")
    _chk(output_height > 0 and output_width > 0, f"{who}: expected output to have non-empty spatial dimensions")
    _chk(input.dtype == th.float32, f"{who}: drtk_b200 computes in float32 only, but input has {input.dtype}; cast it to float32")
    return _GridScatter.apply(input, grid, int(output_height), int(output_width), _PADS[padding_mode], _MODES[mode],
                              bool(align_corners))


class _GridScatterRef(th.autograd.Function):
    """out = d/d tex sum(grid_sample(tex, grid) * input).  torch has no double backward for grid_sample, so the
    backward is spelled out: d/d input = grid_sample(grad_out, grid), d/d grid = grad of sum(that * input)."""

    @staticmethod
    def forward(ctx, input, grid, out_h, out_w, kw):
        tex = th.zeros((input.shape[0], input.shape[1], out_h, out_w), dtype=input.dtype, device=input.device,
                       requires_grad=True)
        with th.enable_grad():
            (out,) = th.autograd.grad(thf.grid_sample(tex, grid.detach(), **kw), tex, grad_outputs=input.detach())
        ctx.save_for_backward(input, grid)
        ctx.kw = kw
        return out

    @staticmethod
    def backward(ctx, grad_out):
        input, grid = ctx.saved_tensors
        g = grid.detach().requires_grad_(True)
        with th.enable_grad():
            sampled = thf.grid_sample(grad_out, g, **ctx.kw)
            (g_grid,) = th.autograd.grad(sampled, g, grad_outputs=input)
        return sampled.detach(), g_grid, None, None, None


def grid_scatter_ref(
    input: th.Tensor,
    grid: th.Tensor,
    output_height: int,
    output_width: int,
    mode: str = "bilinear",
    padding_mode: str = "border",
    align_corners: Optional[bool] = None,
) -> th.Tensor:
    """Stock-torch statement (any device / dtype) built on grid_sample's own backward; counterpart of the
    reference's `grid_scatter_ref` (`drtk/grid_scatter.py:108-191`).  Note: for bicubic sampling with border /
    reflection padding and coordinates outside the image the native op pads the sample position itself
    (`src/grid_scatter/grid_scatter_kernel.cu:141-142`), grid_sample does not -- the two agree inside the image."""
    kw = dict(mode=mode, padding_mode=padding_mode, align_corners=align_corners)
    return _GridScatterRef.apply(input, grid, int(output_height), int(output_width), kw)
