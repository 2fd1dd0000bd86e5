"""transform fwd+bwd: the fused kernels vs the stock-torch-op statement (what the reference runs), config-4 vertex table.
usage: python tools/transform_bench.py [--V 50625] [--N 8]"""
import argparse, os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
T = sys.modules["drtk_b200.transform"]
from tests.util import random_cameras as cameras

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8)
ap.add_argument("--V", type=int, default=50625)
ap.add_argument("--iters", type=int, default=50)
a = ap.parse_args()
dev = "cuda:0"
g = th.Generator().manual_seed(3)
cam = [x.float().to(dev).requires_grad_(True) for x in cameras(a.N, g)]
v = (th.rand((a.N, a.V, 3), generator=g) + th.tensor([-0.5, -0.5, 2.0])).to(dev).requires_grad_(True)
w = th.rand((a.N, a.V, 3), generator=g).to(dev)

def timeit(fn):
    for _ in range(5): fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters): fn()
    e1.record(); th.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters * 1e3

for mode, nd in ((None, 0), ("radial-tangential", 8), ("fisheye", 4), ("fisheye62", 8)):
    D = (th.rand((a.N, nd), generator=g) * 0.02).to(dev).requires_grad_(True) if nd else None
    fov = th.full((a.N, 1), 0.6, device=dev) if nd else None
    def step(fn):
        vp, _ = fn(v, *cam, mode, D, fov)
        vp.backward(w)
    print(f"{str(mode):18s} fused {timeit(lambda: step(T.project_points)):8.1f} us   torch ops {timeit(lambda: step(T.project_points_ref)):8.1f} us  (fwd+bwd, N={a.N}, V={a.V})")
