"""ctypes/numpy front-end of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg -- never by drtk_b200/.  Each function restates the reference
algorithm cited in drtk_oracle_impl.h; see that file for the file:line map.

All arrays are numpy, C-contiguous; `vi` may be [F,3] (shared) or [N,F,3].
dtype float32 -> *_f32 build, float64 -> *_f64 build.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("drtk_oracle.c", "drtk_oracle_impl.h")]
    stale = (not os.path.exists(so)) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _sfx(dt):
    if dt == np.float32:
        return "_f32"
    if dt == np.float64:
        return "_f64"
    raise TypeError(f"oracle: unsupported dtype {dt}")


def _prep_vi(vi, N):
    vi = np.ascontiguousarray(vi, dtype=np.int32)
    if vi.ndim == 2:
        return vi, 0, vi.shape[0]
    assert vi.shape[0] == N
    return vi, 1, vi.shape[1]


def _c(a, dt=None):
    return np.ascontiguousarray(a if dt is None else a.astype(dt, copy=False))


I64 = ctypes.c_int64


def rasterize(v, vi, H, W, mode=0, with_margin=False):
    """-> (depth_img f32 [N,H,W], index_img i32 [N,H,W][, margin u32 [N,H,W]])."""
    v = _c(v)
    N, V, _ = v.shape
    vi, vb, F = _prep_vi(vi, N)
    depth = np.empty((N, H, W), np.float32)
    index = np.empty((N, H, W), np.int32)
    margin = np.empty((N, H, W), np.uint32) if with_margin else None
    fn = getattr(lib(), "oracle_rasterize" + _sfx(v.dtype))
    fn(_p(v), _p(vi), I64(N), I64(V), I64(F), I64(H), I64(W), ctypes.c_int(vb), ctypes.c_int(mode),
       _p(depth), _p(index), _p(margin))
    return (depth, index, margin) if with_margin else (depth, index)


def rasterize_lines(v, vi, H, W):
    """Wireframe mode -> (depth_img f32 [N,H,W], index_img i32 [N,H,W]); edge flags in the top nibble of vi[...,0]."""
    v = _c(v)
    N, V, _ = v.shape
    vi, vb, F = _prep_vi(vi, N)
    depth = np.empty((N, H, W), np.float32)
    index = np.empty((N, H, W), np.int32)
    fn = getattr(lib(), "oracle_rasterize_lines" + _sfx(v.dtype))
    fn(_p(v), _p(vi), I64(N), I64(V), I64(F), I64(H), I64(W), ctypes.c_int(vb), _p(depth), _p(index))
    return depth, index


def render_fwd(v, vi, index_img):
    v = _c(v)
    N, V, _ = v.shape
    vi, vb, F = _prep_vi(vi, N)
    index_img = _c(index_img, np.int32)
    _, H, W = index_img.shape
    depth = np.empty((N, H, W), v.dtype)
    bary = np.empty((N, 3, H, W), v.dtype)
    fn = getattr(lib(), "oracle_render_fwd" + _sfx(v.dtype))
    fn(_p(v), _p(vi), _p(index_img), I64(N), I64(V), I64(F), I64(H), I64(W), ctypes.c_int(vb),
       _p(depth), _p(bary))
    return depth, bary


def render_bwd(v, vi, index_img, grad_depth, grad_bary):
    v = _c(v)
    N, V, _ = v.shape
    vi, vb, F = _prep_vi(vi, N)
    index_img = _c(index_img, np.int32)
    _, H, W = index_img.shape
    gd = None if grad_depth is None else _c(grad_depth, v.dtype)
    gb = None if grad_bary is None else _c(grad_bary, v.dtype)
    grad_v = np.empty((N, V, 3), v.dtype)
    fn = getattr(lib(), "oracle_render_bwd" + _sfx(v.dtype))
    fn(_p(v), _p(vi), _p(index_img), _p(gd), _p(gb), I64(N), I64(V), I64(F), I64(H), I64(W),
       ctypes.c_int(vb), _p(grad_v))
    return grad_v


def interpolate_fwd(attr, vi, index_img, bary_img):
    attr = _c(attr)
    N, V, C = attr.shape
    vi, vb, F = _prep_vi(vi, N)
    index_img = _c(index_img, np.int32)
    bary_img = _c(bary_img, attr.dtype)
    _, H, W = index_img.shape
    out = np.empty((N, C, H, W), attr.dtype)
    fn = getattr(lib(), "oracle_interpolate_fwd" + _sfx(attr.dtype))
    fn(_p(attr), _p(vi), _p(index_img), _p(bary_img), I64(N), I64(V), I64(F), I64(C), I64(H), I64(W),
       ctypes.c_int(vb), _p(out))
    return out


def interpolate_bwd(grad_out, attr, vi, index_img, bary_img, need_attr=True, need_bary=True):
    attr = _c(attr)
    N, V, C = attr.shape
    vi, vb, F = _prep_vi(vi, N)
    index_img = _c(index_img, np.int32)
    bary_img = _c(bary_img, attr.dtype)
    grad_out = _c(grad_out, attr.dtype)
    _, H, W = index_img.shape
    ga = np.empty((N, V, C), attr.dtype) if need_attr else None
    gb = np.empty((N, 3, H, W), attr.dtype) if need_bary else None
    fn = getattr(lib(), "oracle_interpolate_bwd" + _sfx(attr.dtype))
    fn(_p(grad_out), _p(attr), _p(vi), _p(index_img), _p(bary_img), I64(N), I64(V), I64(F), I64(C),
       I64(H), I64(W), ctypes.c_int(vb), _p(ga), _p(gb))
    return ga, gb


def edge_grad_bwd(v_pix, img, index_img, vi, grad_output, max_dp_dr=1e4):
    v_pix = _c(v_pix)
    N, V, _ = v_pix.shape
    vi, vb, F = _prep_vi(vi, N)
    index_img = _c(index_img, np.int32)
    img = _c(img, v_pix.dtype)
    grad_output = _c(grad_output, v_pix.dtype)
    _, C, H, W = img.shape
    out = np.empty((N, 3, H, W), v_pix.dtype)
    real = ctypes.c_float if v_pix.dtype == np.float32 else ctypes.c_double
    fn = getattr(lib(), "oracle_edge_grad_bwd" + _sfx(v_pix.dtype))
    fn(_p(v_pix), _p(img), _p(index_img), _p(vi), _p(grad_output), I64(N), I64(V), I64(F), I64(C),
       I64(H), I64(W), ctypes.c_int(vb), real(max_dp_dr), _p(out))
    return out


# ------------------------------------------------------------------------------------------------
# sparse interpolation matrices (numpy restatement; small cases only)
#   interpolation_matrix         src/interpolate/interpolate_kernel.cu:301-338 (+ launcher :699-762)
#   interpolation_normal_matrix  structure: src/interpolate/interpolate_module.cpp:120-222,
#                                values: src/interpolate/interpolate_kernel.cu:378-416
# ------------------------------------------------------------------------------------------------
def interpolation_matrix(vi, index_img, bary_img, num_vertices):
    """-> dict(crow, col, values, row_pixels, dense): CSR of A [R, V] and its dense form (float64)."""
    index_img = np.asarray(index_img, np.int32)
    N, H, W = index_img.shape
    vi = np.asarray(vi, np.int32)
    if vi.ndim == 2:
        vi = np.broadcast_to(vi[None], (N,) + vi.shape)
    flat = index_img.reshape(-1)
    row_pixels = np.flatnonzero(flat != -1).astype(np.int64)          # :733-735
    R = row_pixels.size
    n = row_pixels // (H * W)
    hw = row_pixels % (H * W)
    cols3 = vi[n, flat[row_pixels]].astype(np.int64)                  # [R,3] corner order
    b3 = np.asarray(bary_img).reshape(N, 3, H * W)[n, :, hw]          # [R,3]
    order = np.argsort(cols3, axis=1, kind="stable")                  # ascending columns (:17-36)
    col = np.take_along_axis(cols3, order, 1).reshape(-1)
    values = np.take_along_axis(b3, order, 1).reshape(-1)
    crow = np.arange(0, 3 * R + 1, 3, dtype=np.int64)
    dense = np.zeros((R, num_vertices), np.float64)
    np.add.at(dense, (np.repeat(np.arange(R), 3), col), values.astype(np.float64))
    return dict(crow=crow, col=col, values=values, row_pixels=row_pixels, dense=dense, order=order)


def interpolation_normal_matrix(vi, index_img, bary_img, num_vertices):
    """-> dict(crow, col, values, pair, dense): CSR of A^T A on the topology's sparsity pattern (values float64)."""
    index_img = np.asarray(index_img, np.int32)
    N, H, W = index_img.shape
    vi = np.asarray(vi, np.int32)
    if vi.ndim == 2:
        vi = np.broadcast_to(vi[None], (N,) + vi.shape)
    v64 = vi.astype(np.int64)
    keys = (v64[:, :, :, None] * num_vertices + v64[:, :, None, :]).reshape(-1)   # module.cpp:166-170
    unique_keys, inverse = np.unique(keys, return_inverse=True)                   # :177-179, :213-217
    rows = unique_keys // max(num_vertices, 1)
    col = unique_keys - rows * num_vertices
    crow = np.zeros(num_vertices + 1, np.int64)
    crow[1:] = np.cumsum(np.bincount(rows, minlength=num_vertices))
    pair = inverse.reshape(N, -1, 9).astype(np.int32)
    values = np.zeros(unique_keys.size, np.float64)
    b = np.asarray(bary_img, np.float64)
    for n in range(N):                                                            # kernel.cu:392-414
        ys, xs = np.nonzero(index_img[n] != -1)
        tri = index_img[n, ys, xs]
        bb = b[n][:, ys, xs]                                                      # [3, P]
        for i in range(3):
            for j in range(3):
                np.add.at(values, pair[n, tri, i * 3 + j], bb[i] * bb[j])
    dense = np.zeros((num_vertices, num_vertices), np.float64)
    dense[rows, col] = values
    return dict(crow=crow, col=col, values=values, pair=pair, dense=dense)


# ---- transform: numpy restatement of the reference's project_points -----------------
def _zsafe(z):
    return np.where(z < 0, np.minimum(z, -1e-8), np.maximum(z, 1e-8))


def _distort(p, mode, D, fov):
    """One batch item.  p [V,2]; D [k]; fov scalar.  Follows drtk/utils/projection.py:
    radial-tangential :56-136, fisheye :139-186, fisheye62 :189-276 (without the LUT)."""
    if mode in (None, "pinhole"):
        return p
    x, y = p[:, 0], p[:, 1]
    if mode == "radial-tangential":
        r2 = np.minimum(x * x + y * y, fov * fov)
        xc, yc = np.clip(x, -fov, fov), np.clip(y, -fov, fov)
        R = 1 + D[0] * r2 + D[1] * r2**2
        if len(D) >= 5:
            R = R + D[4] * r2**3
        if len(D) == 8:
            R = R / (1 + D[5] * r2 + D[6] * r2**2 + D[7] * r2**3)
        q = p * R[:, None]
        q = q + 2 * (xc * yc)[:, None] * np.array([D[2], D[3]])
        q = q + r2[:, None] * np.array([D[3], D[2]])
        q = q + np.stack((2 * D[3] * xc**2, 2 * D[2] * yc**2), -1)
        return q
    r = np.sqrt(x * x + y * y)
    r = np.minimum(np.maximum(r, 1e-8), fov)
    th_ = np.arctan(r)
    nk = 4 if mode == "fisheye" else 6
    thd = th_ * (1 + sum(D[i] * th_ ** (2 * i + 2) for i in range(nk)))
    q = p * (thd / np.maximum(r, 1e-8))[:, None]
    if nk == 6:
        q = np.clip(q, -fov, fov)
        xr, yr = q[:, 0], q[:, 1]
        rr2 = xr * xr + yr * yr
        q = q + np.stack(((2 * xr * xr + rr2) * D[6] + 2 * xr * yr * D[7],
                          2 * xr * yr * D[6] + (2 * yr * yr + rr2) * D[7]), -1)
    return q


def transform_fwd(v, campos, camrot, focal, princpt, modes=None, D=None, fov=None, cull=False):
    """-> (v_pix, v_cam) float64.  `modes`: one string or a list of N strings; `fov` [N] or [N,1] (required for
    the distorted models: the FOV estimators are host logic, tested separately); `cull`: fisheye62 with a
    user-given fov marks vertices beyond it with z = -1 (projection.py:624-644)."""
    v = np.asarray(v, np.float64)
    N = v.shape[0]
    if not isinstance(modes, (list, tuple)):
        modes = [modes] * N
    v_pix, v_cam = np.empty_like(v), np.empty_like(v)
    for n in range(N):
        vc = (v[n] - np.asarray(campos[n], np.float64)) @ np.asarray(camrot[n], np.float64).T  # :536
        z = vc[:, 2:3]
        p = vc[:, :2] / _zsafe(z)
        f = None if fov is None else float(np.asarray(fov, np.float64).reshape(N)[n])
        q = _distort(p, modes[n], None if D is None else np.asarray(D[n], np.float64), f)
        pix = q @ np.asarray(focal[n], np.float64).T + np.asarray(princpt[n], np.float64)
        zc = z
        if cull and modes[n] in ("fisheye62", "fisheye62_lut"):
            zc = np.where(np.sqrt((p**2).sum(-1, keepdims=True)) > f, -1.0, z)
        v_pix[n] = np.concatenate((pix, zc), -1)
        v_cam[n] = vc
    return v_pix, v_cam


def transform_vjp_fd(w_pix, w_cam, v, campos, camrot, focal, princpt, modes=None, D=None, fov=None, cull=False,
                     h=1e-6):
    """Gradients of L = sum(w_pix*v_pix) + sum(w_cam*v_cam) by central differences in float64 (small cases).
    -> dict with v, campos, camrot, focal, princpt, D."""
    args = dict(v=np.asarray(v, np.float64), campos=np.asarray(campos, np.float64),
                camrot=np.asarray(camrot, np.float64), focal=np.asarray(focal, np.float64),
                princpt=np.asarray(princpt, np.float64))
    if D is not None:
        args["D"] = np.asarray(D, np.float64)

    def per_item_loss(a):
        vp, vc = transform_fwd(a["v"], a["campos"], a["camrot"], a["focal"], a["princpt"], modes, a.get("D"), fov, cull)
        return (w_pix * vp).sum((1, 2)) + (w_cam * vc).sum((1, 2)), (w_pix * vp), (w_cam * vc)

    out = {}
    for name, x in args.items():
        g = np.zeros_like(x)
        if name == "v":  # each vertex only feeds its own outputs: perturb one coordinate of all vertices at once
            for k in range(3):
                lo, hi = dict(args), dict(args)
                d = np.zeros_like(x); d[..., k] = h
                hi["v"], lo["v"] = x + d, x - d
                _, ap, ac = per_item_loss(hi)
                _, bp, bc = per_item_loss(lo)
                g[..., k] = ((ap - bp).sum(-1) + (ac - bc).sum(-1)) / (2 * h)
        else:  # per-item parameters: perturb one entry of every item at once
            flat = x.reshape(x.shape[0], -1)
            gf = g.reshape(x.shape[0], -1)
            for k in range(flat.shape[1]):
                d = np.zeros_like(flat); d[:, k] = h
                lo, hi = dict(args), dict(args)
                hi[name], lo[name] = (flat + d).reshape(x.shape), (flat - d).reshape(x.shape)
                gf[:, k] = (per_item_loss(hi)[0] - per_item_loss(lo)[0]) / (2 * h)
        out[name] = g
    return out


# ---- screen_space_uv_derivative: numpy restatement of drtk/screen_space_uv_derivative.py:16-80 -------------
def screen_space_uv_derivative(v, vt, vi, vti, index_img, bary_img, mask, campos, camrot, focal):
    """float64.  v [N,V,3], vt [N,T,2], vi / vti [F,3], index_img [N,H,W], bary_img [N,3,H,W], mask [N,H,W] bool,
    pinhole camera.  -> [N,H,W,2,2] = [[du/dx, dv/dx], [du/dy, dv/dy]], zero outside mask.  Pixels with index -1
    are zero too (the reference evaluates interpolate's background sweep there; callers mask them)."""
    v, vt, bary = (np.asarray(a, np.float64) for a in (v, vt, bary_img))
    N, H, W = index_img.shape
    out = np.zeros((N, H, W, 2, 2))
    for n in range(N):
        live = (index_img[n] >= 0) & np.asarray(mask[n], bool)
        t = index_img[n][live]
        corners, uv = v[n][vi[t]], vt[n][vti[t]]                                   # [P,3,3], [P,3,2]
        dpdb = corners[:, 1:3] - corners[:, 0:1]                                   # geometry.py:73
        dtdb = uv[:, 1:3] - uv[:, 0:1]
        dpdt = np.linalg.solve(dtdb, dpdb)                                         # :77-81, [P,2,3]
        b = bary[n][:, live].T                                                     # [P,3]
        dpdt = dpdt * b.sum(1)[:, None, None]       # interpolating a per-face constant scales it by sum(b)
        p = (corners * b[:, :, None]).sum(1)
        R, c, Fm = (np.asarray(a[n], np.float64) for a in (camrot, campos, focal))
        dcam = dpdt @ R.T                                                          # projection.py:683-684
        pcam = (p - c) @ R.T
        z = pcam[:, 2]
        z = np.where(z < 0, np.minimum(z, -1e-8), np.maximum(z, 1e-8))
        dproj = (dcam[:, :, :2] * z[:, None, None] - pcam[:, None, :2] * dcam[:, :, 2:3]) / (z * z)[:, None, None]
        dpix = dproj @ Fm.T                                                        # [P, i, j] = d pix_j / d t_i
        out[n][live] = np.linalg.inv(dpix)
    return out
