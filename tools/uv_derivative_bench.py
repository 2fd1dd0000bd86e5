"""screen_space_uv_derivative at config-4 size: fused kernel vs the reference-style composition (2 interpolates + torch ops).
usage: python tools/uv_derivative_bench.py"""
import os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
from drtk_b200 import scenes
S = sys.modules["drtk_b200.screen_space_uv_derivative"]
dev = "cuda:0"
v_pix, vi, H, W = scenes.config_mesh(4, N=4, device=dev)
N, V = v_pix.shape[:2]
index = drtk_b200.rasterize(v_pix, vi, H, W)
_, bary = drtk_b200.render(v_pix, vi, index)
v = th.cat(((v_pix[..., :2] - W / 2) / 1000.0 * v_pix[..., 2:], v_pix[..., 2:]), -1).contiguous()
vt = th.rand((N, V, 2), device=dev)
campos, camrot = th.zeros(N, 3, device=dev), th.eye(3, device=dev)[None].expand(N, -1, -1).contiguous()
focal = (th.eye(2, device=dev) * 1000.0)[None].expand(N, -1, -1).contiguous()
mask = index != -1
args = (v, vt, vi, vi, index, bary, mask, campos, camrot, focal)
def timeit(fn, it=10):
    for _ in range(2): fn()
    th.cuda.synchronize()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); th.cuda.synchronize()
    return e0.elapsed_time(e1) / it
with th.no_grad():
    a = drtk_b200.screen_space_uv_derivative(*args)
    b = S._composed(*args, None, None)
    print("max |fused - composed| / scale:", float((a - b).abs().max() / b.abs().max()))
    print(f"N={N} {H}x{W}: fused {timeit(lambda: drtk_b200.screen_space_uv_derivative(*args)):.3f} ms   composition {timeit(lambda: S._composed(*args, None, None), 3):.3f} ms")
