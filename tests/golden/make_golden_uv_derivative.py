#!/usr/bin/env python
"""Generate tests/golden/uvd_*.npz from the UNMODIFIED reference `screen_space_uv_derivative`
(drtk/screen_space_uv_derivative.py), run on CPU in this container: its Python files are loaded straight from
/root/reference and its `interpolate` is the reference's own CPU kernel from oracle/_ref (oracle/ref.py).

    python tests/golden/make_golden_uv_derivative.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch as th

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from drtk_b200 import scenes  # noqa: E402  (pure-torch scene generators)
from oracle import ref as R  # noqa: E402
from tests.util import random_cameras  # noqa: E402

REF = "/root/reference/drtk"


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


pkg = types.ModuleType("drtk"); pkg.__path__ = []
utils = types.ModuleType("drtk.utils"); utils.__path__ = []
sys.modules.update({"drtk": pkg, "drtk.utils": utils})
load("drtk.utils.indexing", f"{REF}/utils/indexing.py")
geo = load("drtk.utils.geometry", f"{REF}/utils/geometry.py")
proj = load("drtk.utils.projection", f"{REF}/utils/projection.py")
utils.face_dpdt, utils.project_points_grad = geo.face_dpdt, proj.project_points_grad
interp = types.ModuleType("drtk.interpolate"); interp.interpolate = R.interpolate
sys.modules["drtk.interpolate"] = interp
SS = load("drtk.screen_space_uv_derivative", f"{REF}/screen_space_uv_derivative.py")

for name, (nx, ny, H, W, seed) in {"uvd_grid_64x80": (9, 8, 64, 80, 11), "uvd_grid_48": (6, 7, 48, 48, 12)}.items():
    N = 2
    g = th.Generator().manual_seed(seed)
    # a mesh in front of a real camera: world-space vertices -> pixels with the reference's own projection
    campos, camrot, focal, princpt = (t.float() for t in random_cameras(N, g))
    focal = focal * (H / 512.0); princpt = th.tensor([[W / 2.0, H / 2.0]]).expand(N, -1)
    v_px, vi = scenes.grid_mesh(nx, ny, H, W, N, seed=seed)
    V = v_px.shape[1]
    z = 2.0 + th.rand((N, V), generator=g)
    xy_cam = (v_px[..., :2] - princpt[:, None]) / th.stack((focal[:, 0, 0], focal[:, 1, 1]), -1)[:, None] * z[..., None]
    v_cam = th.cat((xy_cam, z[..., None]), -1)
    v = th.einsum("nji,nvj->nvi", camrot, v_cam) + campos[:, None]           # world = R^T cam + c
    v_pix, _ = proj.project_points(v, campos, camrot, focal, princpt)
    index = R.rasterize(v_pix, vi, H, W)
    _, bary = R.render(v_pix, vi, index)
    vt = th.rand((N, V + 3, 2), generator=g)
    vti = (vi.long() + th.randint(0, 3, (1,), generator=g)).int()
    mask = index != -1
    out = SS.screen_space_uv_derivative(v, vt, vi, vti, index, bary, mask, campos, camrot, focal)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), v=v.numpy(), vt=vt.numpy(), vi=vi.numpy(), vti=vti.numpy(),
                        index_img=index.numpy(), bary_img=bary.numpy(), mask=mask.numpy(), campos=campos.numpy(),
                        camrot=camrot.numpy(), focal=focal.numpy(), out=out.numpy())
    print(name, float(out.abs().max()), float(mask.float().mean()))
