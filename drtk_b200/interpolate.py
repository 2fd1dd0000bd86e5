"""drtk.interpolate on the B200 kernels (API mirror of the reference `drtk/interpolate.py:19-50`).

Autograd contract of the reference's InterpolateFunction
(`src/interpolate/interpolate_module.cpp:378-425`): gradients to `vert_attributes` and/or
`bary_img` according to their requires_grad; an undefined upstream gradient yields no gradients.
"""
import torch as th

from . import _ops, torch_ops


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
        native = (th.float32, th.float64)
        a32 = attr.detach() if attr.dtype in native else attr.detach().float()
        b32 = bary_img.detach() if bary_img.dtype in native else bary_img.detach().float()
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
    if torch_ops.enabled():
        return torch_ops.interpolate(vert_attributes, vi, index_img, bary_img)
    vert_attributes, bary_img = _ops.autocast_f32(vert_attributes, bary_img)
    return _InterpolateFn.apply(vert_attributes, vi, index_img, bary_img)


# ------------------------------------------------------------------------------------------------
# sparse interpolation matrices (API mirror of the reference `drtk/interpolate.py:53-192`)
# ------------------------------------------------------------------------------------------------
class _InterpolationMatrixFn(th.autograd.Function):
    """InterpolationMatrixFunction (`src/interpolate/interpolate_module.cpp:435-475`): the CSR indices and the row ->
    pixel map are discrete (non-differentiable); the values are the barycentrics, so their gradient flows to bary_img."""

    @staticmethod
    def forward(ctx, vi, index_img, bary_img):
        crow, col, values, row_pixels = _ops.interpolation_matrix_forward(vi, index_img, bary_img)
        ctx.save_for_backward(vi, index_img, row_pixels)
        ctx.bary_dtype = bary_img.dtype
        ctx.set_materialize_grads(False)
        ctx.mark_non_differentiable(crow, col, row_pixels)
        return crow, col, values.to(bary_img.dtype), row_pixels

    @staticmethod
    def backward(ctx, g_crow, g_col, g_values, g_rows):
        if g_values is None or not ctx.needs_input_grad[2]:
            return None, None, None
        vi, index_img, row_pixels = ctx.saved_tensors
        return None, None, _ops.interpolation_matrix_backward(g_values, vi, index_img, row_pixels).to(ctx.bary_dtype)


class _NormalMatrixValuesFn(th.autograd.Function):
    """Values of A^T A and their product-rule gradient w.r.t. bary_img
    (`src/interpolate/interpolate_module.cpp:486-560`, kernels `src/interpolate/interpolate_kernel.cu:378-452`)."""

    @staticmethod
    def forward(ctx, pair_indices, index_img, bary_img, nnz):
        values = _ops.interpolation_normal_matrix_values(pair_indices, index_img, bary_img, nnz)
        ctx.save_for_backward(pair_indices, index_img, bary_img)
        ctx.set_materialize_grads(False)
        return values.to(bary_img.dtype)

    @staticmethod
    def backward(ctx, g_values):
        if g_values is None or not ctx.needs_input_grad[2]:
            return None, None, None, None
        pair_indices, index_img, bary_img = ctx.saved_tensors
        gb = _ops.interpolation_normal_matrix_values_backward(g_values, pair_indices, index_img, bary_img.detach())
        return None, None, gb.to(bary_img.dtype), None


def _broadcast_vi(vi, n):
    if vi.ndim == 2:
        return vi[None].expand(n, -1, -1)
    if vi.ndim == 3 and vi.shape[0] == 1 and n != 1:
        return vi.expand(n, -1, -1)
    return vi


@th.compiler.disable
def interpolation_matrix(vi: th.Tensor, index_img: th.Tensor, bary_img: th.Tensor, num_vertices: int) -> th.Tensor:
    """Sparse CSR matrix A [num_valid_pixels, num_vertices] with `pixel_values = A @ X` for per-vertex attributes X:
    one row per foreground pixel (flattened [N,H,W] order, background skipped), three entries per row = the pixel's
    barycentrics at the triangle's vertex columns, columns ascending within a row.  Gradients flow to bary_img
    through the values.  (Reference `drtk/interpolate.py:53-126`.)"""
    vi = _broadcast_vi(vi, index_img.shape[0])
    crow, col, values, row_pixels = _InterpolationMatrixFn.apply(vi, index_img, bary_img)
    return th.sparse_csr_tensor(crow, col, values, size=(int(row_pixels.numel()), int(num_vertices)),
                                device=values.device, dtype=values.dtype, check_invariants=False)


# topology cache of interpolation_normal_matrix: CSR structure + per-face pair lookup, keyed like the reference's
# NormalMatrixCacheKey (`src/interpolate/interpolate_module.cpp:36-118`): identity + version of the vi tensor, so
# that iterative solvers that keep the same topology tensor never rebuild (or synchronise) again.
import collections
import threading

_NM_CACHE_MAX = 128
_nm_cache = collections.OrderedDict()
_nm_lock = threading.Lock()


def _nm_key(vi, num_vertices, device):
    return (str(device), str(vi.device), vi.untyped_storage().data_ptr(), vi.data_ptr(), tuple(vi.shape), tuple(vi.stride()),
            vi.storage_offset(), str(vi.dtype), int(num_vertices), vi._version)


def _build_normal_matrix_structure(vi, num_vertices, device):
    """crow_indices [V+1] i64, col_indices [nnz] i64, pair_indices [N,F,9] i32
    (`src/interpolate/interpolate_module.cpp:120-222`: keys row*V+col of the nine directed vertex pairs of every
    face, sorted unique keys -> CSR, lower_bound of every key -> pair slot).  Built with torch ops on the CPU like
    the reference (a cache miss copies vi to the host)."""
    if num_vertices < 0:
        raise RuntimeError("interpolation_normal_matrix(): expected num_vertices to be non-negative")
    if num_vertices > 2 ** 31 - 1:
        raise RuntimeError("interpolation_normal_matrix(): expected num_vertices to fit in int32")
    vi_cpu = vi.detach().to("cpu", th.int32).contiguous().to(th.int64)
    N, F = vi_cpu.shape[0], vi_cpu.shape[1]
    if N * F > 0 and num_vertices <= 0:
        raise RuntimeError("interpolation_normal_matrix(): expected num_vertices to be positive when faces are present")
    if N * F > 0 and (int(vi_cpu.min()) < 0 or int(vi_cpu.max()) >= num_vertices):
        raise RuntimeError("interpolation_normal_matrix(): vi contains a vertex index outside [0, num_vertices)")
    keys = (vi_cpu[:, :, :, None] * num_vertices + vi_cpu[:, :, None, :]).reshape(-1)  # [N*F*9], (i, j) -> i*3+j
    unique_keys, inverse = th.unique(keys, sorted=True, return_inverse=True)
    if unique_keys.numel() > 2 ** 31 - 1:
        raise RuntimeError("interpolation_normal_matrix(): normal matrix has too many nonzeros for int32 value indices")
    rows = th.div(unique_keys, max(num_vertices, 1), rounding_mode="floor")
    col = unique_keys - rows * num_vertices
    crow = th.zeros(num_vertices + 1, dtype=th.int64)
    if unique_keys.numel():
        crow[1:] = th.cumsum(th.bincount(rows, minlength=num_vertices), 0)
    pair = inverse.to(th.int32).reshape(N, F, 9)
    return crow.to(device), col.to(device), pair.to(device)


def _normal_matrix_structure(vi, num_vertices, device):
    key = _nm_key(vi, num_vertices, device)
    with _nm_lock:
        hit = _nm_cache.get(key)
        if hit is not None:
            _nm_cache.move_to_end(key)
            return hit[1]
    structure = _build_normal_matrix_structure(vi, num_vertices, device)  # outside the lock, like the reference
    with _nm_lock:
        hit = _nm_cache.get(key)
        if hit is not None:
            return hit[1]
        while len(_nm_cache) >= _NM_CACHE_MAX:
            _nm_cache.popitem(last=False)
        _nm_cache[key] = (vi, structure)  # keeps vi alive so its pointers cannot be recycled into a stale hit
    return structure


@th.compiler.disable
def interpolation_normal_matrix(vi: th.Tensor, index_img: th.Tensor, bary_img: th.Tensor, num_vertices: int) -> th.Tensor:
    """Sparse CSR normal matrix A^T A [num_vertices, num_vertices] of :func:`interpolation_matrix`, assembled
    directly: every foreground pixel adds the nine products bary_i * bary_j at the entries (vi[i], vi[j]) of its
    triangle.  The sparsity pattern depends on the topology only and is cached per vi tensor (identity + version).
    Differentiable w.r.t. bary_img.  (Reference `drtk/interpolate.py:129-192`.)"""
    vi = _broadcast_vi(vi, index_img.shape[0])
    _ops._check_matrix_inputs("interpolation_normal_matrix", vi, index_img, bary_img)
    crow, col, pair = _normal_matrix_structure(vi, int(num_vertices), bary_img.device)
    values = _NormalMatrixValuesFn.apply(pair, index_img, bary_img, int(col.numel()))
    return th.sparse_csr_tensor(crow, col, values, size=(int(num_vertices), int(num_vertices)),
                                device=values.device, dtype=values.dtype, check_invariants=False)


def interpolate_ref(vert_attributes: th.Tensor, vi: th.Tensor, index_img: th.Tensor, bary_img: th.Tensor) -> th.Tensor:
    """Pure-PyTorch, float64, differentiable statement of :func:`interpolate` (any device), for tests and debugging
    -- the counterpart of the reference's `interpolate_ref` (`drtk/interpolate.py:195-261`): same values, same
    coordinate sweep on empty pixels.  `vi` is [F,3]."""
    dt = vert_attributes.dtype
    a, b = vert_attributes.double(), bary_img.double()
    N, H, W = index_img.shape
    C = a.shape[-1]
    tri = index_img.clamp(min=0).long()                       # [N,H,W]
    corners = vi.long()[tri]                                  # [N,H,W,3]
    rows = th.gather(a, 1, corners.reshape(N, -1, 1).expand(-1, -1, C)).reshape(N, H, W, 3, C)
    img = (rows * b.permute(0, 2, 3, 1)[..., None]).sum(3)    # [N,H,W,C]
    xs = (th.arange(W, device=a.device, dtype=th.float64) * 2 + 1) / W - 1
    ys = (th.arange(H, device=a.device, dtype=th.float64) * 2 + 1) / H - 1
    sweep = th.stack((xs[None, :].expand(H, W), ys[:, None].expand(H, W)), -1)        # even channels x, odd y
    sweep = sweep.repeat(1, 1, (C + 1) // 2)[..., :C]
    img = th.where((index_img == -1)[..., None], sweep[None].expand(N, -1, -1, -1), img)
    return img.permute(0, 3, 1, 2).to(dt)
