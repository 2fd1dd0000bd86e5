"""drtk.transform: world -> pixel projection (pure PyTorch plumbing, CPU or GPU tensors).

API mirror of the reference `drtk/transform.py:14-119` for the undistorted pinhole camera
(`drtk/utils/projection.py:33-53`, `:536`).  This step sits in front of the hot path, is O(V)
and autograd-differentiable through stock torch ops; it needs no kernel.  Lens-distortion modes
of the reference (radial-tangential, fisheye, fisheye62) are outside the accelerated path and
are not provided here.
"""
from typing import List, Optional, Tuple, Union

import torch as th


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
) -> Tuple[th.Tensor, th.Tensor]:
    """Returns (v_pix, v_cam), both [N,V,3]; v_pix = (x_pixels, y_pixels, z_camera)."""
    if not ((camrot is not None and campos is not None) ^ (Rt is not None)):
        raise ValueError("You must provide exactly one of Rt or (campos, camrot).")
    if not ((focal is not None and princpt is not None) ^ (K is not None)):
        raise ValueError("You must provide exactly one of K or (focal, princpt).")
    modes = distortion_mode if isinstance(distortion_mode, (list, tuple)) else [distortion_mode]
    if any(m not in (None, "pinhole") for m in modes):
        raise NotImplementedError(
            f"drtk_b200.transform: distortion mode {distortion_mode!r} is not provided; only the "
            "pinhole camera is (lens distortion is outside the accelerated path)")
    if Rt is not None:
        camrot = Rt[:, :3, :3]
        campos = -(camrot.transpose(-2, -1) @ Rt[:, :3, 3:4])[..., 0]
    if K is not None:
        focal = K[:, :2, :2]
        princpt = K[:, :2, 2]
    # v_cam = R (v - c)
    v_cam = th.einsum("nij,nvj->nvi", camrot, v - campos[:, None])
    z = v_cam[..., 2:3]
    # keep |z| >= 1e-8 with its sign so the perspective divide is finite
    z_safe = th.where(z < 0, z.clamp(max=-1e-8), z.clamp(min=1e-8))
    xy = th.einsum("nij,nvj->nvi", focal, v_cam[..., :2] / z_safe) + princpt[:, None]
    return th.cat((xy, z), dim=-1), v_cam


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
