"""mipmap_grid_sample / grid_scatter fwd+bwd: drtk_b200 vs the reference CUDA kernels (oracle/_ref) on a render-like
workload: N images of HxW pixels look up an SxS RGB texture pyramid through a smooth uv field (rotation + zoom +
low-frequency warp), Jacobian taken analytically from the same field.
usage: python tools/sampler_bench.py [--N 8] [--H 2048] [--S 1024] [--C 3] [--aniso 4] [--iters 10]"""
import argparse, math, os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
from oracle import ref as R

ap = argparse.ArgumentParser()
ap.add_argument("--N", type=int, default=8)
ap.add_argument("--H", type=int, default=2048)
ap.add_argument("--S", type=int, default=1024)
ap.add_argument("--C", type=int, default=3)
ap.add_argument("--aniso", type=int, default=4)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--mode", default="bilinear")
a = ap.parse_args()
dev = "cuda:0"
N, H, W, S, C = a.N, a.H, a.H, a.S, a.C


def uv_field():
    """uv in [-1,1] (grid_sample convention) and d(uv01)/d(pixel) [N,H,W,2,2] (rows: d/dx, d/dy; uv01 = (uv+1)/2)."""
    ys, xs = th.meshgrid(th.arange(H, device=dev, dtype=th.float32), th.arange(W, device=dev, dtype=th.float32), indexing="ij")
    x, y = (xs + 0.5) / W * 2 - 1, (ys + 0.5) / H * 2 - 1
    grids, jacs = [], []
    for n in range(N):
        ang, zoom = 0.15 * n, 0.6 + 0.1 * n
        c, s = math.cos(ang) * zoom, math.sin(ang) * zoom
        u = c * x - s * y + 0.05 * th.sin(3 * y + n)
        v = s * x + c * y + 0.05 * th.cos(2 * x - n)
        dudx = (c + 0 * x) * (2 / W); dudy = (-s + 0.15 * th.cos(3 * y + n)) * (2 / H)
        dvdx = (s - 0.10 * th.sin(2 * x - n)) * (2 / W); dvdy = (c + 0 * y) * (2 / H)
        grids.append(th.stack((u, v), -1))
        jacs.append(th.stack((th.stack((dudx, dvdx), -1), th.stack((dudy, dvdy), -1)), -2) * 0.5)
    return th.stack(grids), th.stack(jacs)


grid, jac = uv_field()
g = th.Generator(device=dev).manual_seed(0)
levels, s = [], S
while s >= 1:
    levels.append(th.rand((N, C, s, s), device=dev, generator=g)); s //= 2
w = th.rand((N, C, H, W), device=dev, generator=g)
img = th.rand((N, C, H, W), device=dev, generator=g)
wt = th.rand((N, C, S, S), device=dev, generator=g)


def timeit(fn):
    for _ in range(3): fn()
    th.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); th.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def mip(fn, grad):
    lv = [t.requires_grad_(grad) for t in levels]
    gr = grid.requires_grad_(grad)
    out = fn(lv, gr, jac, a.aniso, a.mode, "border", False)
    if grad:
        out.backward(w)
        for t in lv: t.grad = None
        gr.grad = None


def scat(fn, grad):
    x, gr = img.requires_grad_(grad), grid.requires_grad_(grad)
    out = fn(x, gr, S, S, a.mode, "border", False)
    if grad:
        out.backward(wt); x.grad = None; gr.grad = None


have_ref = R.samplers_available()
mpix = N * H * W / 1e6
print(f"N={N} {H}x{W} px, texture {S}^2 x{len(levels)} levels, C={C}, max_aniso={a.aniso}, {a.mode}")
for name, run, mine, ref in (("mipmap_grid_sample", mip, drtk_b200.mipmap_grid_sample, R.mipmap_grid_sample),
                             ("grid_scatter", scat, drtk_b200.grid_scatter, R.grid_scatter)):
    for grad in (False, True):
        t_new = timeit(lambda: run(mine, grad))
        t_ref = timeit(lambda: run(ref, grad)) if have_ref else float("nan")
        print(f"{name:20s} {'fwd+bwd' if grad else 'fwd    '}  drtk_b200 {t_new:8.3f} ms ({mpix / t_new * 1e3:8.0f} Mpix/s)   reference CUDA {t_ref:8.3f} ms   x{t_ref / t_new:.2f}")
