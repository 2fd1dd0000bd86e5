"""Two steps of the config-4 pipeline for ncu (profiles/): first step warms up, second is captured.
usage: ncu ... python tools/profile_step.py [--config 4] [--steps 2] [--ref]"""
import argparse, os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
from drtk_b200 import scenes
ap = argparse.ArgumentParser(); ap.add_argument("--config", type=int, default=4); ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--ref", action="store_true"); ap.add_argument("--N", type=int, default=None)
a = ap.parse_args()
dev = "cuda:0"
v, vi, H, W = scenes.config_mesh(a.config, N=a.N, device=dev)
N = v.shape[0]
attr = scenes.vertex_attributes(N, v.shape[1], 16, seed=1, device=dev)
w = th.rand((N, 16, H, W), device=dev)
api = drtk_b200
if a.ref:
    from oracle import ref as api
for _ in range(a.steps):
    vv, aa = v.clone().requires_grad_(True), attr.clone().requires_grad_(True)
    index = api.rasterize(vv, vi, H, W)
    _, bary = api.render(vv, vi, index)
    img = api.interpolate(aa, vi, index, bary)
    img = api.edge_grad_estimator(vv, vi, bary, img, index)
    img.backward(gradient=w)
th.cuda.synchronize()
print("done")
