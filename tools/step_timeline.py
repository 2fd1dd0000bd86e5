"""Where does a config-4 step go?  GPU kernel list (torch profiler) + CPU issue time vs GPU time.
usage: python tools/step_timeline.py"""
import os, sys, time
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
from drtk_b200 import scenes
dev = "cuda:0"
v, vi, H, W = scenes.config_mesh(4, device=dev)
N = v.shape[0]
attr = scenes.vertex_attributes(N, v.shape[1], 16, seed=1, device=dev)
w = th.rand((N, 16, H, W), device=dev)
v.requires_grad_(True); attr.requires_grad_(True)

class WS(th.autograd.Function):
    @staticmethod
    def forward(ctx, img, w):
        ctx.save_for_backward(w); return (img * w).sum()
    @staticmethod
    def backward(ctx, g):
        return ctx.saved_tensors[0] * g, None

def step():
    v.grad = None; attr.grad = None
    index = drtk_b200.rasterize(v, vi, H, W)
    _, bary = drtk_b200.render(v, vi, index)
    img = drtk_b200.interpolate(attr, vi, index, bary)
    img = drtk_b200.edge_grad_estimator(v, vi, bary, img, index)
    loss = WS.apply(img, w)
    loss.backward()
    return loss

for _ in range(5): step()
th.cuda.synchronize()
K = 20
e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for _ in range(K): step()
t1 = time.perf_counter(); e1.record(); th.cuda.synchronize()
print(f"cpu issue {1e3*(t1-t0)/K:.3f} ms/step   gpu {e0.elapsed_time(e1)/K:.3f} ms/step")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): step()
    th.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3, e.count / 3) for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
rows.sort(key=lambda r: -r[1])
tot = 0
for k, t, c in rows:
    print(f"{t:9.1f} us  x{c:4.1f}  {k[:110]}"); tot += t
print(f"sum of device time {tot/1e3:.3f} ms/step")
