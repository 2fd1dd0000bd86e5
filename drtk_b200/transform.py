"""drtk.transform: world -> pixel projection of the vertex table.

API mirror of the reference `drtk/transform.py:14-119` (`transform`, `transform_with_v_cam`) and of
`drtk/utils/projection.py:486-646` (`project_points`), all camera models included: pinhole,
"radial-tangential", "fisheye", "fisheye62" / "fisheye62_lut", and per-batch lists of the first three.

CUDA tensors go through ONE kernel per direction (`csrc/transform.cu`, C ABI `drtk_b200_transform_forward /
_backward`) instead of the reference's ~20 stock torch kernels forward and as many again backward; the
gradients of `campos / camrot / focal / princpt / distortion_coeff` come back through the packed camera block
the host builds with a single `torch.cat`.  CPU tensors (BASELINE config 1, "plumbing, no GPU") take the
pure-torch statement `project_points_ref` below -- that is the reference's own situation (it has no native
code on this step), not a fallback of a CUDA op: every op with a kernel in the reference still raises on CPU.

Not differentiated: `fov` (the reference estimates it under `no_grad` when it is not given).
"""
from typing import List, Optional, Tuple, Union

import numpy as np
import torch as th

from . import _lib

_MODE_ID = {None: 0, "pinhole": 0, "radial-tangential": 1, "fisheye": 2, "fisheye62": 3, "fisheye62_lut": 3}
_LIST_MODES = (None, "pinhole", "radial-tangential", "fisheye")  # projection.py:12-17
_CAM = 28


# ---- field-of-view bounds (host, numpy root finding; not differentiable) -----------------------------------
def _smallest_positive_root(desc_coefs: np.ndarray) -> Optional[float]:
    roots = np.roots(desc_coefs)
    real = roots.real[np.abs(roots.imag) < 1e-5]
    real = real[real > 0]
    return float(real.min()) if real.size else None


def _as_numpy(D) -> np.ndarray:
    return D.detach().cpu().numpy() if th.is_tensor(D) else np.asarray(D)


def _like(fov: np.ndarray, D):
    out = np.asarray(fov, dtype=np.float32)[:, None]
    return th.from_numpy(out).to(D) if th.is_tensor(D) else out


def estimate_rt_fov(D) -> th.Tensor:
    """Largest normalised radius up to which r*(1 + k1 r^2 + k2 r^4) is monotonic: the smallest positive root of
    its derivative 1 + 3 k1 r^2 + 5 k2 r^4, inf when there is none (`projection.py:279-327`).  -> [N,1]"""
    d = _as_numpy(D)
    fov = []
    for k in d:
        r = _smallest_positive_root(np.array([5 * k[1], 0, 3 * k[0], 0, 1], dtype=d.dtype))
        fov.append(np.inf if r is None else r)
    return _like(np.asarray(fov), D)


def _fisheye_fov(d: np.ndarray, nk: int, D):
    fov = []
    for k in d:
        desc = []
        for i in range(nk - 1, -1, -1):  # derivative of theta + sum k_i theta^(2i+3), descending powers
            desc += [(2 * i + 3) * k[i], 0]
        r = _smallest_positive_root(np.array(desc + [1], dtype=d.dtype))
        fov.append(np.pi / 2 if r is None else min(r, np.pi / 2))
    return _like(np.tan(np.asarray(fov)), D)


def estimate_fisheye_fov(D) -> th.Tensor:
    """tan of the first positive angle (capped at pi/2) where the 4-coefficient fisheye polynomial stops being
    monotonic (`projection.py:358-399`).  -> [N,1]"""
    return _fisheye_fov(_as_numpy(D), 4, D)


def estimate_fisheye62_fov(D) -> th.Tensor:
    """Same with the six radial coefficients of fisheye62 (`projection.py:402-457`)."""
    d = _as_numpy(D)
    assert d.shape[-1] >= 6, f"fisheye62 FOV requires at least 6 coefficients, got shape {d.shape}"
    return _fisheye_fov(d, 6, D)


# ---- pure-torch statement (CPU tensors; also what the tests differentiate in float64) ----------------------
def _zsafe(z):
    return th.where(z < 0, z.clamp(max=-1e-8), z.clamp(min=1e-8))


def _distort_ref(p, mode, D, fov):
    """p [N,V,2] normalised image plane -> distorted plane; `mode` one of the single-model strings."""
    if mode in (None, "pinhole"):
        return p
    x, y = p[..., 0], p[..., 1]
    if mode == "radial-tangential":
        assert D.shape[1] in (4, 5, 8)
        r2 = (x * x + y * y).clamp(max=fov.pow(2))
        xc, yc = x.clamp(min=-fov, max=fov), y.clamp(min=-fov, max=fov)
        radial = 1 + D[:, 0:1] * r2 + D[:, 1:2] * r2.pow(2)
        if D.shape[1] >= 5:
            radial = radial + D[:, 4:5] * r2.pow(3)
        if D.shape[1] == 8:
            radial = radial / (1 + D[:, 5:6] * r2 + D[:, 6:7] * r2.pow(2) + D[:, 7:8] * r2.pow(3))
        p1, p2 = D[:, 2:3], D[:, 3:4]
        qx = x * radial + 2 * xc * yc * p1 + r2 * p2 + 2 * p2 * xc * xc
        qy = y * radial + 2 * xc * yc * p2 + r2 * p1 + 2 * p1 * yc * yc
        return th.stack((qx, qy), -1)
    nk = 4 if mode == "fisheye" else 6
    r = (x * x + y * y).sqrt()
    rc = r.clamp(min=1e-8 * th.ones_like(fov), max=fov)
    theta = th.atan(rc)
    poly = 1
    for i in range(nk):
        poly = poly + D[:, i:i + 1] * theta.pow(2 * i + 2)
    q = p * (theta * poly / rc.clamp(min=1e-8))[..., None]
    if nk == 6:
        q = q.clamp(min=-fov[..., None], max=fov[..., None])
        xr, yr = q[..., 0], q[..., 1]
        rr2 = xr * xr + yr * yr
        p0, p1 = D[:, 6:7], D[:, 7:8]
        q = q + th.stack(((2 * xr * xr + rr2) * p0 + 2 * xr * yr * p1,
                          2 * xr * yr * p0 + (2 * yr * yr + rr2) * p1), -1)
    return q


def _default_fov(mode, D):
    with th.no_grad():
        return estimate_rt_fov(D) if mode == "radial-tangential" else estimate_fisheye_fov(D)


def _lut_offset(pix, lut_vector_field, lut_spacing):
    """fisheye62_lut: bilinear lookup of a pixel-space correction, zero outside the table
    (`projection.py:245-276`)."""
    assert lut_spacing is not None, "lookup table spacing must be provided along with vector field"
    g = pix / lut_spacing[:, None, :]
    cols, rows = lut_vector_field.shape[2:4]
    g = th.stack((g[..., 0] / (cols - 1) * 2.0 - 1.0, g[..., 1] / (rows - 1) * 2.0 - 1.0), -1)
    off = th.nn.functional.grid_sample(lut_vector_field, g[:, None], align_corners=True)[:, :, 0].transpose(1, 2)
    outside = (g.abs() > 1.0).any(-1, keepdim=True)
    return th.where(outside, th.zeros_like(off), off)


def project_points_ref(v, campos, camrot, focal, princpt, distortion_mode=None, distortion_coeff=None, fov=None,
                       lut_vector_field=None, lut_spacing=None) -> Tuple[th.Tensor, th.Tensor]:
    """Stock torch ops, any device / dtype.  Same contract as `project_points`."""
    mode, modes = _resolve_modes(distortion_mode, distortion_coeff)
    v_cam = th.einsum("nij,nvj->nvi", camrot, v - campos[:, None])
    z = v_cam[..., 2:3]
    p = v_cam[..., :2] / _zsafe(z)
    fov_given = fov is not None
    if modes is None:
        if mode not in (None, "pinhole") and fov is None:
            fov = _default_fov(mode, distortion_coeff)
        q = _distort_ref(p, mode, distortion_coeff, fov)
    else:
        q = th.empty_like(p)
        for m in set(modes):
            sel = th.tensor([x == m for x in modes], device=v.device)
            if m in (None, "pinhole"):
                q[sel] = p[sel]
                continue
            f = fov[sel] if fov is not None else _default_fov(m, distortion_coeff[sel])
            q[sel] = _distort_ref(p[sel], m, distortion_coeff[sel], f)
    pix = th.einsum("nij,nvj->nvi", focal, q) + princpt[:, None]
    if mode in ("fisheye62", "fisheye62_lut"):
        if lut_vector_field is not None:
            pix = pix + _lut_offset(pix, lut_vector_field, lut_spacing)
        if fov_given:
            z = th.where(p.pow(2).sum(-1, keepdim=True).sqrt() > fov.view(-1, 1, 1), th.full_like(z, -1.0), z)
    return th.cat((pix, z), -1), v_cam


# ---- the CUDA path -----------------------------------------------------------------------------------------
def _resolve_modes(distortion_mode, distortion_coeff):
    """-> (single mode string or None, per-item list or None); errors as `projection.py:533-597`."""
    if distortion_mode is not None:
        assert distortion_coeff is not None, "Missing distortion coefficients."
    if isinstance(distortion_mode, (list, tuple)):
        uniq = set(distortion_mode)
        if len(uniq) == 0:
            return None, None
        if len(uniq) == 1:
            distortion_mode = next(iter(uniq))
        else:
            if not uniq <= set(_LIST_MODES):
                raise ValueError(f"Invalid distortion mode: {distortion_mode}. Valid options: {set(_LIST_MODES)}.")
            return None, list(distortion_mode)
    if distortion_mode is not None and not isinstance(distortion_mode, str) or distortion_mode not in _MODE_ID:
        raise ValueError(f"Invalid distortion mode: {distortion_mode}. Valid options: {set(_LIST_MODES)}.")
    return distortion_mode, None


class _Transform(th.autograd.Function):
    """(v [N,V,3], cam [N,28]) -> (v_pix, v_cam); one kernel forward, one backward."""

    @staticmethod
    def forward(ctx, v, cam, modes_dev, mode_id, cull):
        lib = _lib.load()
        N, V = v.shape[0], v.shape[1]
        with th.cuda.device(v.device):
            v_pix = th.empty((N, V, 3), dtype=th.float32, device=v.device)
            v_cam = th.empty((N, V, 3), dtype=th.float32, device=v.device)
            rc = lib.drtk_b200_transform_forward(
                _lib.ptr(v), _lib.strides(v), _lib.ptr(cam), _lib.ptr(modes_dev), mode_id, int(cull), N, V,
                _lib.ptr(v_pix), _lib.ptr(v_cam), th.cuda.current_stream(v.device).cuda_stream)
        _lib.check(rc, "transform()")
        ctx.save_for_backward(v, cam, modes_dev)
        ctx.mode_id, ctx.cull = mode_id, cull
        return v_pix, v_cam

    @staticmethod
    def backward(ctx, g_pix, g_cam):
        v, cam, modes_dev = ctx.saved_tensors
        need_v, need_cam = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        if (g_pix is None and g_cam is None) or not (need_v or need_cam):
            return None, None, None, None, None
        lib = _lib.load()
        N, V = v.shape[0], v.shape[1]
        g_pix = None if g_pix is None else g_pix.float()
        g_cam = None if g_cam is None else g_cam.float()
        with th.cuda.device(v.device):
            grad_v = th.empty((N, V, 3), dtype=th.float32, device=v.device) if need_v else None
            grad_cam = th.empty((N, _CAM), dtype=th.float32, device=v.device) if need_cam else None
            rc = lib.drtk_b200_transform_backward(
                _lib.ptr(v), _lib.strides(v), _lib.ptr(cam), _lib.ptr(modes_dev), ctx.mode_id, int(ctx.cull),
                _lib.ptr(g_pix), None if g_pix is None else _lib.strides(g_pix),
                _lib.ptr(g_cam), None if g_cam is None else _lib.strides(g_cam), N, V,
                _lib.ptr(grad_v), _lib.ptr(grad_cam), th.cuda.current_stream(v.device).cuda_stream)
        _lib.check(rc, "transform() backward")
        return grad_v, grad_cam, None, None, None


def _f32(t, name):
    if t.dtype == th.float32:
        return t
    if t.dtype in (th.float16, th.bfloat16):
        return t.float()
    raise RuntimeError(f"transform(): drtk_b200 computes in float32 only, but {name} has {t.dtype}; cast it to float32")


@th.compiler.disable  # ctypes launch inside: keep torch.compile out, like the reference's native ops
def project_points(v, campos, camrot, focal, princpt, distortion_mode=None, distortion_coeff=None, fov=None,
                   lut_vector_field=None, lut_spacing=None) -> Tuple[th.Tensor, th.Tensor]:
    """-> (v_pix, v_cam), both [N,V,3]; v_pix = (x_pixels, y_pixels, z_camera).  `drtk/utils/projection.py:486-646`."""
    if not v.is_cuda or v.dtype == th.float64:
        # CPU tensors (no kernel in the reference either) and float64 (the reference's torch-op chain works in any dtype;
        # the fused kernel is float32): stock torch ops
        return project_points_ref(v, campos, camrot, focal, princpt, distortion_mode, distortion_coeff, fov,
                                  lut_vector_field, lut_spacing)
    mode, modes = _resolve_modes(distortion_mode, distortion_coeff)
    N = v.shape[0]
    v = _f32(v, "v")
    D = distortion_coeff
    fov_given = fov is not None
    if modes is None:
        if mode == "radial-tangential":
            assert D.shape[1] in (4, 5, 8)
        if mode in ("fisheye62", "fisheye62_lut"):
            assert D.shape[1] == 8, f"Fisheye62 model requires 8 distortion parameters: {D.shape}"
        if mode not in (None, "pinhole") and fov is None:
            fov = _default_fov(mode, D)
        modes_dev = None
    else:
        if fov is None:  # per item, by its own model (what the reference's sub-batch calls do)
            fov = th.ones((N, 1), dtype=th.float32, device=v.device)
            for m in set(modes) - {None, "pinhole"}:
                sel = th.tensor([x == m for x in modes], device=v.device)
                fov[sel] = _default_fov(m, D[sel]).float()
        modes_dev = th.tensor([_MODE_ID[m] for m in modes], dtype=th.int32, device=v.device)
    parts = [_f32(campos, "campos").reshape(N, 3), _f32(camrot, "camrot").reshape(N, 9),
             _f32(focal, "focal").reshape(N, 4), _f32(princpt, "princpt").reshape(N, 2)]
    nd = 0
    if D is not None and (modes is not None or mode not in (None, "pinhole")):
        nd = D.shape[1]
        parts.append(_f32(D, "distortion_coeff").reshape(N, nd))
    tail = th.zeros((N, _CAM - 18 - nd), dtype=th.float32, device=v.device)
    if fov is not None:
        tail[:, 26 - 18 - nd] = fov.detach().reshape(N).float()
    cam = th.cat(parts + [tail], 1)  # the one torch op on the way in; autograd splits the block's gradient
    v_pix, v_cam = _Transform.apply(v, cam, modes_dev, _MODE_ID[mode] if modes is None else 0, fov_given)
    if mode in ("fisheye62", "fisheye62_lut") and lut_vector_field is not None:
        v_pix = th.cat((v_pix[..., :2] + _lut_offset(v_pix[..., :2], lut_vector_field, lut_spacing),
                        v_pix[..., 2:]), -1)
    return v_pix, v_cam


def transform_with_v_cam(
    v: th.Tensor,
    campos: Optional[th.Tensor] = None,
    camrot: Optional[th.Tensor] = None,
    focal: Optional[th.Tensor] = None,
    princpt: Optional[th.Tensor] = None,
    K: Optional[th.Tensor] = None,
    Rt: Optional[th.Tensor] = None,
    distortion_mode: Optional[Union[List[str], str]] = None,
    distortion_coeff: Optional[th.Tensor] = None,
    fov: Optional[th.Tensor] = None,
    lut_vector_field: Optional[th.Tensor] = None,
    lut_spacing: Optional[th.Tensor] = None,
) -> Tuple[th.Tensor, th.Tensor]:
    """Returns (v_pix, v_cam), both [N,V,3]  (`drtk/transform.py:68-119`)."""
    if not ((camrot is not None and campos is not None) ^ (Rt is not None)):
        raise ValueError("You must provide exactly one of Rt or (campos, camrot).")
    if not ((focal is not None and princpt is not None) ^ (K is not None)):
        raise ValueError("You must provide exactly one of K or (focal, princpt).")
    if Rt is not None:
        camrot = Rt[:, :3, :3]
        campos = -(camrot.transpose(-2, -1) @ Rt[:, :3, 3:4])[..., 0]
    if K is not None:
        focal = K[:, :2, :2]
        princpt = K[:, :2, 2]
    return project_points(v, campos, camrot, focal, princpt, distortion_mode, distortion_coeff, fov,
                          lut_vector_field, lut_spacing)


def transform(
    v: th.Tensor,
    campos: Optional[th.Tensor] = None,
    camrot: Optional[th.Tensor] = None,
    focal: Optional[th.Tensor] = None,
    princpt: Optional[th.Tensor] = None,
    K: Optional[th.Tensor] = None,
    Rt: Optional[th.Tensor] = None,
    distortion_mode: Optional[Union[List[str], str]] = None,
    distortion_coeff: Optional[th.Tensor] = None,
    fov: Optional[th.Tensor] = None,
) -> th.Tensor:
    """Project vertices v [N,V,3] to the image plane; returns [N,V,3] = (x, y, z_cam).

    Provide either K [N,3,3] or (focal [N,2,2], princpt [N,2]); and either Rt [N,3,4]/[N,4,4]
    or (campos [N,3], camrot [N,3,3]).  With Rt = [R|t]: camrot = R, campos = -R^T t.
    """
    return transform_with_v_cam(v, campos, camrot, focal, princpt, K, Rt, distortion_mode,
                                distortion_coeff, fov)[0]
