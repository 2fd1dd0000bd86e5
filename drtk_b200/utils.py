"""The slice of `drtk.utils` that sits on the projection step in front of the hot path
(`drtk/utils/__init__.py`, `drtk/utils/projection.py`): `project_points` and its helpers.  The geometry
utilities of the reference's `drtk.utils` (normals, tangent frames, index images, ...) are outside the path.
"""
from typing import List, Optional, Tuple, Union

import torch as th

from .transform import (  # noqa: F401
    estimate_fisheye62_fov,
    estimate_fisheye_fov,
    estimate_rt_fov,
    project_points,
    project_points_ref,
)


def project_points_grad(
    v_grad: th.Tensor,
    v: th.Tensor,
    campos: th.Tensor,
    camrot: th.Tensor,
    focal: th.Tensor,
    distortion_mode: Optional[Union[List[str], str]] = None,
    distortion_coeff: Optional[th.Tensor] = None,
) -> th.Tensor:
    """Jacobian-vector product of the pinhole projection: pixel-space perturbation [N,V,2] caused by the
    world-space perturbation `v_grad` [N,V,3] (`drtk/utils/projection.py:649-706`; like the reference, the
    distorted models are not implemented here).  Used by `screen_space_uv_derivative`."""
    if distortion_mode is not None:
        assert distortion_coeff is not None, "Missing distortion coefficients."
        if distortion_mode in ("radial-tangential", "fisheye"):
            raise NotImplementedError
        raise ValueError(f"Invalid distortion mode: {distortion_mode}.")
    d_cam = th.einsum("nij,nvj->nvi", camrot, v_grad)
    v_cam = th.einsum("nij,nvj->nvi", camrot, v - campos[:, None])
    z = v_cam[..., 2:3]
    z = th.where(z < 0, z.clamp(max=-1e-8), z.clamp(min=1e-8))
    # quotient rule on v_cam.xy / z
    d_proj = (d_cam[..., :2] * z - v_cam[..., :2] * d_cam[..., 2:3]) / (z * z)
    return th.einsum("nij,nvj->nvi", focal, d_proj)


def face_dpdt(v: th.Tensor, vt: th.Tensor, vi: th.Tensor, vti: th.Tensor) -> Tuple[th.Tensor, th.Tensor]:
    """Per-triangle transposed Jacobian of the 3-D position w.r.t. the uv coordinates (`drtk/utils/geometry.py:18-82`):
    (dp/dt)^T = ((dt/db)^T)^-1 (dp/db)^T with b the barycentrics.  v [N,V,3], vt [N,T,2], vi / vti [F,3] (int64).
    Returns (dpdt_t [N,F,2,3] with [..., i, j] = dp_j/dt_i, and the per-face vertex positions [N,F,3,3])."""
    if v.ndim != 3:
        raise ValueError(f"Expected v to be 3D, got {v.ndim}D")
    if vt.ndim != 3:
        raise ValueError(f"Expected vt to be 3D, got {vt.ndim}D")
    if vt.shape[0] != v.shape[0]:
        raise ValueError(f"Expected vt to have the same batch size as v, got {vt.shape[0]} and {v.shape[0]}")
    corners, uv = v[:, vi], vt[:, vti]
    edges_p = corners[:, :, 1:3] - corners[:, :, 0:1]
    edges_t = uv[:, :, 1:3] - uv[:, :, 0:1]
    return th.linalg.solve(edges_t, edges_p), corners
