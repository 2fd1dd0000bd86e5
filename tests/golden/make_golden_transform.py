#!/usr/bin/env python
"""Generate tests/golden/transform_*.npz from the UNMODIFIED reference `project_points`.

    python tests/golden/make_golden_transform.py        (build container: /root/reference must exist)

The reference's projection is pure Python/torch (drtk/utils/projection.py); the file is loaded straight from
/root/reference, run on CPU in float64 (values and autograd gradients of a seeded linear loss), and inputs +
outputs are stored so that the oracle and the CUDA kernel can be held to them where the reference is absent.
"""
import importlib.util
import os
import sys

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.util import random_cameras as cameras  # noqa: E402
spec = importlib.util.spec_from_file_location("ref_projection", "/root/reference/drtk/utils/projection.py")
P = importlib.util.module_from_spec(spec)
spec.loader.exec_module(P)


CASES = {
    # name: (mode, D columns, D scale, give fov)
    "pinhole": (None, 0, 0.0, False),
    "rt4": ("radial-tangential", 4, 0.05, False),
    "rt5_fov": ("radial-tangential", 5, 0.05, True),
    "rt8_fov": ("radial-tangential", 8, 0.05, True),
    "fisheye": ("fisheye", 4, 0.02, False),
    "fisheye_fov": ("fisheye", 4, 0.02, True),
    "fisheye62": ("fisheye62", 8, 0.01, False),
    "fisheye62_fov": ("fisheye62", 8, 0.01, True),
    "mixed": (["pinhole", "radial-tangential", "fisheye"], 4, 0.03, False),
}


def main():
    N, V = 3, 257
    for i, (name, (mode, nd, dscale, give_fov)) in enumerate(CASES.items()):
        g = th.Generator().manual_seed(4200 + i)
        campos, camrot, focal, princpt = cameras(N, g)
        # a cloud in front of the camera, wide enough that some points pass the fov clamps; vertex 0 sits behind
        # the camera
        v = th.rand((N, V, 3), generator=g, dtype=th.float64) * th.tensor([4.0, 4.0, 2.0], dtype=th.float64) + th.tensor([-2.0, -2.0, 1.0], dtype=th.float64)
        v[:, 0, 2] = -0.5
        D = (th.rand((N, nd), generator=g, dtype=th.float64) * 2 - 1) * dscale if nd else None
        fov = (0.6 + 0.3 * th.rand((N, 1), generator=g, dtype=th.float64)) if give_fov else None
        w_pix = th.rand((N, V, 3), generator=g, dtype=th.float64)
        w_cam = th.rand((N, V, 3), generator=g, dtype=th.float64)
        leaves = [t.clone().requires_grad_(True) for t in (v, campos, camrot, focal, princpt)]
        Dl = D.clone().requires_grad_(True) if D is not None else None
        v_pix, v_cam = P.project_points(*leaves, distortion_mode=mode, distortion_coeff=Dl, fov=fov)
        loss = (v_pix * w_pix).sum() + (v_cam * w_cam).sum()
        grads = th.autograd.grad(loss, leaves + ([Dl] if Dl is not None else []), allow_unused=True)
        # the fov the reference used (it estimates one when none is given)
        if fov is not None or mode is None:
            fov_used = fov
        elif isinstance(mode, list):
            fov_used = th.ones((N, 1), dtype=th.float64)
            for n, m in enumerate(mode):
                if m == "radial-tangential":
                    fov_used[n] = P.estimate_rt_fov(D[n:n + 1])[0]
                elif m == "fisheye":
                    fov_used[n] = P.estimate_fisheye_fov(D[n:n + 1])[0]
        else:
            fov_used = P.estimate_rt_fov(D) if mode == "radial-tangential" else P.estimate_fisheye_fov(D)
        out = dict(v=v, campos=campos, camrot=camrot, focal=focal, princpt=princpt, w_pix=w_pix, w_cam=w_cam,
                   v_pix=v_pix.detach(), v_cam=v_cam.detach(), g_v=grads[0], g_campos=grads[1], g_camrot=grads[2],
                   g_focal=grads[3], g_princpt=grads[4])
        if D is not None:
            out.update(D=D, g_D=grads[5] if grads[5] is not None else th.zeros_like(D))
        if fov is not None:
            out["fov"] = fov
        if fov_used is not None:
            out["fov_used"] = fov_used
        arrays = {k: t.numpy() for k, t in out.items()}
        arrays["mode"] = np.array(mode if isinstance(mode, list) else [str(mode)])
        np.savez_compressed(os.path.join(HERE, f"transform_{name}.npz"), **arrays)
        print(name, "ok", float(v_pix.abs().max()))


if __name__ == "__main__":
    main()
