"""The slice of `drtk.utils` that sits on the projection step in front of the hot path
(`drtk/utils/__init__.py`, `drtk/utils/projection.py`): `project_points` and its helpers.  The geometry
utilities of the reference's `drtk.utils` (normals, tangent frames, index images, ...) are outside the path.
"""
from typing import List, Optional, Union

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
