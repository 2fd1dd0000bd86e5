"""A/B timing of single ops on config-4 tensors (CUDA events, L2-cold by construction: each op streams >1 GB).
usage: python tools/opbench.py [--ops interp_bwd,...] [--iters 20]
The library reads its DRTK_B200_* developer switches ONCE per process, so an A/B is two invocations:
    python tools/opbench.py --ops interp_bwd --dump /tmp/a.pt
    DRTK_B200_MERGED=1 python tools/opbench.py --ops interp_bwd --cmp /tmp/a.pt"""
import argparse, os, sys
import torch as th
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drtk_b200
from drtk_b200 import scenes, _ops

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, default=4)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--C", type=int, default=16)
ap.add_argument("--ops", default="interp_bwd")
ap.add_argument("--dump", default="", help="save each op's outputs to this file")
ap.add_argument("--cmp", default="", help="compare each op's outputs with a file written by --dump")
ap.add_argument("--overdraw", action="store_true")
a = ap.parse_args()
dev = "cuda:0"
v, vi, H, W = scenes.config_mesh(a.config, overdraw=a.overdraw, device=dev)
N = v.shape[0]
vi3 = vi[None].expand(N, -1, -1)
attr = scenes.vertex_attributes(N, v.shape[1], a.C, seed=1, device=dev)
w = th.rand((N, a.C, H, W), device=dev, generator=th.Generator(device=dev).manual_seed(2))  # seeded: --dump / --cmp runs see the same cotangent
depth, index = _ops.rasterize(v, vi3, H, W)
_, bary = _ops.render_forward(v, vi3, index)
img = _ops.interpolate_forward(attr, vi3, index, bary)

def timeit(fn):
    for _ in range(3): fn()
    th.cuda.synchronize()
    evs = []
    for _ in range(a.iters):
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); evs.append((e0, e1))
    th.cuda.synchronize()
    ts = sorted(x.elapsed_time(y) for x, y in evs)
    return ts[len(ts) // 2], ts[0], out

vi_wire = vi3.clone()
vi_wire[..., 0] |= (7 << 28)  # all edges visible (wireframe mode, src/rasterize/rasterize_kernel.cu:293-303)


def _ref_op(name, *args):
    from oracle import ref as R
    R.load()
    return getattr(getattr(th.ops, name + "_ext"), {"rasterize": "rasterize", "render": "render", "interpolate": "interpolate"}[name])(*args)


OPS = {
    "rasterize": lambda: _ops.rasterize(v, vi3, H, W),
    "wireframe": lambda: _ops.rasterize(v, vi_wire, H, W, wireframe=True),
    "wireframe_ref": lambda: tuple(_ref_op("rasterize", v, vi_wire, H, W, True)),
    "rasterize_ref": lambda: tuple(_ref_op("rasterize", v, vi3, H, W, False)),
    "render_fwd": lambda: _ops.render_forward(v, vi3, index),
    "interp_fwd": lambda: _ops.interpolate_forward(attr, vi3, index, bary),
    "interp_bwd": lambda: _ops.interpolate_backward(w, attr, vi3, index, bary, True, True),
    "interp_bwd_v": lambda: _ops.interpolate_backward(w, attr, vi3, index, bary, True, False),
    "interp_bwd_b": lambda: _ops.interpolate_backward(w, attr, vi3, index, bary, False, True),
    "render_bwd": lambda: _ops.render_backward(v, vi3, index, None, bary),
    "edge_fused": lambda: _ops.edge_grad_backward_fused(v, img, index, vi3, w, bary, 1e4),
}
if any(o.startswith("nm_") or o.startswith("imat") for o in a.ops.split(",")):
    import importlib
    I = importlib.import_module("drtk_b200.interpolate")
    crow, col, pair = I._normal_matrix_structure(vi3, v.shape[1], th.device(dev))
    nnz = int(col.numel())
    gvals = th.rand((nnz,), device=dev)
    OPS["nm_values"] = lambda: _ops.interpolation_normal_matrix_values(pair, index, bary, nnz)
    OPS["nm_values_bwd"] = lambda: _ops.interpolation_normal_matrix_values_backward(gvals, pair, index, bary)
    OPS["imat"] = lambda: _ops.interpolation_matrix_forward(vi3, index, bary)[2]
    try:
        from oracle import ref as R
        R.load()
        vic = vi3.contiguous()
        OPS["nm_values_ref"] = lambda: th.ops.interpolate_ext.interpolation_normal_matrix_values(pair, index, bary, nnz)
        OPS["imat_ref"] = lambda: th.ops.interpolate_ext.interpolation_matrix(vic, index, bary)[2]
    except Exception as ex:  # noqa: BLE001
        print("reference ops unavailable:", ex)
tag = ",".join(f"{k}={v_}" for k, v_ in sorted(os.environ.items()) if k.startswith("DRTK_B200_")) or "default"
saved = th.load(a.cmp) if a.cmp else {}
dump = {}
for op in a.ops.split(","):
    med, mn, out = timeit(OPS[op])
    outs = [o for o in (out if isinstance(out, tuple) else (out,)) if o is not None]
    msg = ""
    for i, (x, y) in enumerate(zip(outs, saved.get(op, []))):
        y = y.to(dev)
        d = (x.float() - y.float()).abs().max().item(); sc = y.float().abs().max().item()
        msg += f" out{i}: maxdiff {d:.3e} (scale {sc:.3e}, finite {bool(th.isfinite(x.float()).all())})"
    if a.dump:
        dump[op] = [o.cpu() for o in outs]
    print(f"{op:14s} {tag:28s} median {med:.4f} ms  min {mn:.4f} ms{msg}", flush=True)
if a.dump:
    th.save(dump, a.dump)
