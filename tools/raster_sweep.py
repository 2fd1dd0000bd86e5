"""Triangle-size sweep of the rasteriser beside the reference CUDA kernel (fill mode), same tensors, bit-compared.
usage: python tools/raster_sweep.py [--H 2048] [--N 8]      (needs oracle/_ref: the reference extension built by build())
Jittered grid meshes over the central 90 % of the canvas whose cells have legs of ~10 / 50 / 200 / 1000 pixels."""
import argparse, os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
from drtk_b200 import scenes, _ops

ap = argparse.ArgumentParser()
ap.add_argument("--H", type=int, default=2048)
ap.add_argument("--N", type=int, default=8)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
dev = "cuda:0"
try:
    from oracle import ref as R
    R.load()
    have_ref = True
except Exception as ex:  # noqa: BLE001
    print("reference extension unavailable:", ex)
    have_ref = False


def timeit(fn, iters):
    for _ in range(2): fn()
    th.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); th.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2], out


print(f"| triangle leg (px) | triangles | drtk_b200 ms | reference kernel ms | ratio | index / depth bits equal |")
print("|---|---|---|---|---|---|")
for leg in (10, 50, 200, 1000):
    nx = max(2, round(0.9 * a.H / leg) + 1)
    v, vi = scenes.grid_mesh(nx, nx, a.H, a.H, a.N, seed=77, device=dev)
    vi3 = vi[None].expand(a.N, -1, -1).contiguous()
    ms, (depth, index) = timeit(lambda: _ops.rasterize(v, vi3, a.H, a.H), a.iters)
    if have_ref:
        ms_r, out_r = timeit(lambda: th.ops.rasterize_ext.rasterize(v, vi3, a.H, a.H, False), max(2, a.iters // 3))
        same = bool((out_r[1] == index).all()) and bool((out_r[0].view(th.int32) == depth.view(th.int32)).all())
        print(f"| {leg} | {vi.shape[0]} | {ms:.3f} | {ms_r:.3f} | {ms_r / ms:.1f}x | {same} |", flush=True)
    else:
        print(f"| {leg} | {vi.shape[0]} | {ms:.3f} | - | - | - |", flush=True)
