"""Guarded first contact of a rasteriser build with the GPU (run under `timeout`): the tiled path (algo 0) must equal
the triangle-parallel validation path (algo 1: same sample arithmetic, global 64-bit atomicMin) bit for bit on small,
odd-sized, overdrawn, large-triangle and BASELINE-size scenes.  Prints per-scene timings; exit code 1 on a mismatch."""
import os, sys, time
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drtk_b200 import _ops, scenes

dev = "cuda:0"
ok = True


def check(name, v, vi, H, W, iters=3):
    global ok
    v, vi = v.to(dev), vi.to(dev)
    vi3 = vi[None].expand(v.shape[0], -1, -1) if vi.ndim == 2 else vi
    d1, i1 = _ops.rasterize(v, vi3, H, W, algo=1)
    d0, i0 = _ops.rasterize(v, vi3, H, W, algo=0)
    th.cuda.synchronize()
    same = bool(th.equal(i0, i1)) and bool(th.equal(d0.view(th.int32), d1.view(th.int32)))
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        _ops.rasterize(v, vi3, H, W, algo=0)
    e1.record()
    th.cuda.synchronize()
    nbad = int((i0 != i1).sum()) + int((d0.view(th.int32) != d1.view(th.int32)).sum())
    print(f"{name:34s} {'OK ' if same else 'BAD'} mismatches {nbad:8d}  covered {int((i0 >= 0).sum()):10d}  {e0.elapsed_time(e1) / iters:8.3f} ms", flush=True)
    ok = ok and same


t0 = time.time()
check("tiny 3x5", *scenes.grid_mesh(3, 3, 3, 5, 1, seed=17), 3, 5)
check("grid 256", *scenes.grid_mesh(21, 21, 256, 256, 2, seed=7), 256, 256)
check("overdraw 192x160", *scenes.grid_mesh(15, 15, 192, 160, 2, seed=11, overdraw=True), 192, 160)
check("odd 97x61 overdraw", *scenes.grid_mesh(9, 7, 61, 97, 3, seed=13, overdraw=True), 61, 97)
g = th.Generator().manual_seed(23)
v = th.rand((2, 30, 3), generator=g) * th.tensor([400.0, 300.0, 3.0]) + th.tensor([-60.0, -40.0, 0.5])
vi = th.randint(0, 30, (24, 3), generator=g, dtype=th.int32)
check("big overlapping tris 280x200", v, vi, 200, 280)
# deep overlap: 40 large triangles over one another (more than four owners per pixel in one pass)
v = th.rand((1, 120, 3), generator=g) * th.tensor([300.0, 300.0, 3.0]) + th.tensor([-20.0, -20.0, 0.5])
vi = th.arange(120, dtype=th.int32).view(40, 3)
check("deep overlap 256", v, vi, 256, 256)
# many tiny triangles in one tile (several passes per tile): 60x60 grid squeezed into 40 px
vv, vii = scenes.grid_mesh(60, 60, 64, 64, 1, seed=5)
vv[..., :2] = vv[..., :2] * 0.6 + 10
check("7k triangles in 40 px", vv, vii, 64, 64)
# integer vertex coordinates: exact ties on edges and vertices
vv, vii = scenes.grid_mesh(17, 17, 128, 128, 1, seed=9)
vv[..., :2] = vv[..., :2].round()
check("integer coordinates 128", vv, vii, 128, 128)
v2, vi2, H2, W2 = scenes.two_triangles()
check("two triangles 512 (config 2)", v2, vi2, H2, W2)
if "--small" in sys.argv:  # under compute-sanitizer: the small scenes only
    print("elapsed", round(time.time() - t0, 1), "s;", "ALL OK" if ok else "FAILED")
    sys.exit(0 if ok else 1)
for cfg, N, od in ((3, 8, False), (3, 2, True), (4, 8, False), (4, 2, True), (5, 1, False)):
    v, vi, H, W = scenes.config_mesh(cfg, N=N, overdraw=od)
    check(f"config {cfg} N={N} overdraw={od}", v, vi, H, W)
# coarse mesh at high resolution: every triangle is "large" (the O(tiles x L) cliff of the round-1 design)
for nx, res in ((32, 2048), (71, 4096)):
    v, vi = scenes.grid_mesh(nx, nx, res, res, 1, seed=3)
    check(f"coarse {2 * (nx - 1) ** 2} tris at {res}^2", v, vi, res, res)
print("elapsed", round(time.time() - t0, 1), "s;", "ALL OK" if ok else "FAILED")
sys.exit(0 if ok else 1)
