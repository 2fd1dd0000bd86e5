#!/usr/bin/env python
"""bench.py -- the DRTK rasterisation hot path on B200: Mpixels/s, forward + backward.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl new|reference] [--config 2|3|4|5] [--overdraw]
                    [--regions R] [--transport auto|nccl|multimem] [--bg-ctas B] [--dispatch auto|torch|ctypes]
                    [--no-extras] [--no-cpu-baseline] [--no-ref-cuda]

One step = one pass of the hot path over one batch of synthetic input (BASELINE.json):

    index = rasterize(v_pix, vi, H, W); depth, bary = render(v_pix, vi, index)
    img = interpolate(attr, vi, index, bary); img = edge_grad_estimator(v_pix, vi, bary, img, index)
    img.backward(gradient=w)                          # v_pix and attr require grad; w = dL/dimg of a linear loss

Default workload: BASELINE config 4 -- 100 352 triangles, 2048x2048, batch 8, 16 vertex
attributes (the configuration the metric is quoted on).  Prints ONE JSON line (rank 0).

  value          whole-job Mpix/s with inputs resident in HBM, CUDA-event timed, max over ranks
  e2e            same metric through the public API with HOST (pinned) inputs: every step copies
                 v_pix / attr to the device (the topology vi stays resident) and reads the step's result
                 back (the gradients summed over the batch and the ranks), copies double-buffered
  roofline       the dominant kernel of the step: algorithmic bytes / its event-timed duration,
                 against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline   the reference's own CPU kernels (oracle/_ref, built from the unmodified reference
                 sources) on a bounded sample of the same workload, on the host cores of the box
  --impl reference   times that CPU implementation as its own arm (rank 0 only under torchrun)

The mesh and the attribute table are parameters SHARED by the batch items, so the result of a step is the gradient
summed over the batch: [V,3] + [V,C].  N > 1 (torchrun, one process per GPU): the batch is sharded -- every rank runs
the same per-GPU workload on its own items (weak scaling, no data-path collective) -- and those sums are all-reduced,
each parameter as soon as autograd has finished it (drtk_b200.dist.SharedGradReducer: one fused batch-sum +
multimem all-reduce kernel over NVSwitch multicast memory, or batch-sum kernel + NCCL).

The K-step region is timed R times (--regions, default 5); `ms_per_step` / `value` are the MEDIAN region, all regions
are listed.  `parity_check` compares the tensors of one step with the reference CUDA kernels run on the same inputs.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch as th

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from drtk_b200 import scenes  # noqa: E402

METRIC = "Mpixels/s fwd+bwd (rasterize+render+interpolate+edge_grad), 100k-tri@2048^2 b8"
METRIC_OTHER = "Mpixels/s fwd+bwd (rasterize+render+interpolate+edge_grad), BASELINE config {cfg}{od}"
C_ATTR = 16

# algorithmic (compulsory) HBM bytes per pixel of each of OUR ops at fp32, C = attribute channels;
# vertex / index tables are L2 resident and excluded (DESIGN.md "Kernels")
ALGO_BYTES_PER_PX = {
    "rasterize": lambda C: 8,                       # index + depth written
    "render_fwd": lambda C: 20,                     # index read; depth + 3 bary written
    "interpolate_fwd": lambda C: 16 + 4 * C,        # index + bary read; C planes written
    # edge_grad: img / grad_out (8C B/px in the reference, which reads them at every triangle change) are
    # only read for pairs that contribute -- data dependent, not compulsory.  The bench pipeline registers
    # no hook, so the fused kernel runs; the two-kernel plan (hook) costs 16 + 28 B/px instead.
    "edge_grad_bwd_fused": lambda C: 4,             # index read; the sparse non-zero gradients go straight to
                                                    # grad_v_pix (REDs into an L2-resident [N,V,3] table)
    "interpolate_bwd": lambda C: 16 + 4 * C + 12,   # index + bary + C grad planes read; bary grad written
    "render_bwd": lambda C: 4 + 12,                 # index + grad_bary read (grad_depth undefined here)
}
# CUDA kernels launched by libdrtk_b200.so per op call (memsets are driver operations, not counted)
KERNELS = {"rasterize": 4, "render_fwd": 1, "interpolate_fwd": 1, "edge_grad_bwd_fused": 2,
           "interpolate_bwd": 2, "render_bwd": 3}
REDUCE_KERNELS = 2  # batch sum (+ in-kernel all-reduce) of grad_v and grad_attr


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed regions (B200_PROFILING.md recipe), through NVML
    in a background thread (every ~2 ms; an nvidia-smi subprocess takes longer to start than the 10-step
    timed region lasts).  Falls back to one nvidia-smi query when pynvml is unavailable."""

    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.rows, self.stop_flag, self.h, self.index = [], False, None, index
        self.windows = []
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                try:
                    phys = int(vis.split(",")[index])
                except Exception:
                    phys = index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.t = threading.Thread(target=self._loop, daemon=True)
            self.t.start()
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((time.perf_counter(), sm, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        if self.h is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=10).stdout.split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1, "reasons": [],
                        "note": "pynvml unavailable: one nvidia-smi sample after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml and nvidia-smi unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=1)
        inside = [r for r in self.rows if any(a <= r[0] <= b for a, b in self.windows)] or self.rows
        sm = [r[1] for r in inside]
        mask = 0
        for r in inside:
            mask |= r[2]
        reasons = sorted(nm for bit, nm in self.REASONS.items() if mask & bit)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": float(self.max_sm), "samples": len(sm),
                "reasons": reasons}


def config_of(cfg):
    """(geometry dict, attribute channels).  Config 2 = the reference demo's two literal triangles (test/two_triangles.py:
    21-32), 512x512, batch 1, 3 colour channels; configs 3-5 = the jittered grid meshes of scenes.CONFIGS, 16 channels."""
    if cfg == 2:
        return dict(H=512, W=512, N=1, F=2), 3
    c = dict(scenes.CONFIGS[cfg])
    c["F"] = 2 * (c["nx"] - 1) * (c["ny"] - 1)
    return c, C_ATTR


def make_inputs(cfg, seed_offset=0, overdraw=False, N=None):
    c, C = config_of(cfg)
    n = c["N"] if N is None else N
    if cfg == 2:
        v, vi, _, _ = scenes.two_triangles()
        v = v.expand(n, -1, -1).contiguous()
    else:
        v, vi = scenes.grid_mesh(c["nx"], c["ny"], c["H"], c["W"], n, seed=1000 * cfg + seed_offset, overdraw=overdraw)
    attr = scenes.vertex_attributes(n, v.shape[1], C, seed=1000 * cfg + 1 + seed_offset)
    return v, vi, attr, c, C


def pipeline(api, v_pix, vi, attr, w, H, W):
    index = api.rasterize(v_pix, vi, H, W)
    _, bary = api.render(v_pix, vi, index)
    img = api.interpolate(attr, vi, index, bary)
    img = api.edge_grad_estimator(v_pix, vi, bary, img, index)
    # backward seeded with the cotangent w, i.e. the gradient of the linear loss (img * w).sum() without
    # materialising it: torch's own elementwise / reduction kernels run at ~1.6 TB/s on this box (img * w:
    # 0.92 ms, .sum(): 0.33 ms, the broadcast mul of the backward: 1.3 ms -- tools/step_timeline.py), which
    # would put 2.5 ms of glue next to 3.3 ms of rasterisation kernels in BOTH arms and blur the comparison.
    img.backward(gradient=w)
    return img


# ------------------------------------------------------------------------------------------------
def run_reference_cpu(cfg, steps, warmup, sample_items=1, overdraw=False):
    """The reference's CPU implementation of the path (oracle/_ref when present, else the oracle
    port) on a bounded sample: `sample_items` batch items of the workload per step."""
    v, vi, attr, c, C = make_inputs(cfg, overdraw=overdraw, N=sample_items)
    w = th.rand((sample_items, C, c["H"], c["W"]), generator=th.Generator().manual_seed(1000 * cfg + 2))
    from oracle import ref as R
    kind = "reference" if R.available() else "port"
    cores = th.get_num_threads()
    if kind == "reference":
        def step():
            vv, aa = v.clone().requires_grad_(True), attr.clone().requires_grad_(True)
            pipeline(R, vv, vi, aa, w, c["H"], c["W"])
    else:
        from oracle import oracle as O
        cores = os.cpu_count() or 1
        vn, vin, an, wn = v.numpy(), vi.numpy(), attr.numpy(), w.numpy()

        def step():
            _, idx = O.rasterize(vn, vin, c["H"], c["W"], mode=0)
            _, bary = O.render_fwd(vn, vin, idx)
            img = O.interpolate_fwd(an, vin, idx, bary)
            gpix = O.edge_grad_bwd(vn, img, idx, vin, wn, 1e4)
            O.interpolate_bwd(gpix, vn, vin, idx, bary, True, False)
            _, gb = O.interpolate_bwd(wn, an, vin, idx, bary, True, True)
            O.render_bwd(vn, vin, idx, None, gb)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    px = sample_items * c["H"] * c["W"]
    return {"value": px / dt / 1e6, "unit": "Mpix/s", "cores": cores, "kind": kind,
            "sample": f"{sample_items} of {c['N']} batch items of config {cfg} per step ({px / 1e6:.2f} Mpix), {steps} steps, {warmup} warm-up",
            "ms_per_step": dt * 1e3, "transform_only": cpu_transform_only(cfg),
            "same_rate_metric": "Mpix/s is a rate: the sample is a fraction of the batch of the same workload, not the whole step"}


def cpu_transform_only(cfg, reps=5):
    """BASELINE's 'transform-only CPU path': the reference's `drtk.transform` is a chain of stock torch ops on the
    [N,V,3] vertex table (drtk/transform.py:68-119 -> drtk/utils/projection.py:33-53, :536), forward + backward on CPU
    tensors of the configuration's size, all host threads torch uses.  Where /root/reference exists (the build
    container) the reference's own file is imported and timed (`kind: reference`); on the GPU box, where it does not,
    the same chain as stated in `drtk_b200.transform.project_points_ref` (`kind: port`)."""
    try:
        import drtk_b200  # noqa: F401  (the package attribute `transform` is the function; the module holds the statement)
        kind, project = "port", sys.modules["drtk_b200.transform"].project_points_ref
        ref_file = "/root/reference/drtk/utils/projection.py"
        if os.path.exists(ref_file):
            try:
                import importlib.util
                spec = importlib.util.spec_from_file_location("_ref_projection", ref_file)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                kind, project = "reference", mod.project_points
            except Exception:  # noqa: BLE001
                pass
        v, _, _, c, _ = make_inputs(cfg)
        N = v.shape[0]
        v = (v + th.tensor([0.0, 0.0, 1.0])).requires_grad_(True)
        cam = (th.zeros(N, 3), th.eye(3)[None].expand(N, -1, -1).contiguous(),
               (th.eye(2) * 1000.0)[None].expand(N, -1, -1).contiguous(), th.full((N, 2), c["W"] / 2.0))
        g = th.ones_like(v)

        def once():
            v.grad = None
            project(v, *cam)[0].backward(g)
        once()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        dt = (time.perf_counter() - t0) / reps
        return {"ms_fwd_bwd": round(dt * 1e3, 3), "Mvertices_per_s": round(v.shape[0] * v.shape[1] / dt / 1e6, 2),
                "cores": th.get_num_threads(), "kind": kind, "vertices": v.shape[0] * v.shape[1]}
    except Exception as ex:  # noqa: BLE001
        return {"unavailable": repr(ex)[:160]}


def compare_with_reference(api_new, api_ref, v, vi, attr, w, H, W):
    """One step of both arms on the same tensors, compared on the device: index_img equal; depth / bary / img within
    1e-5 (relative + the same fraction of the tensor's scale); vertex / attribute gradients within 5e-5 of scale (both
    sides are long atomically ordered fp32 sums).  NaN / Inf anywhere fails.  Returns the `parity_check` object."""
    res = {}
    for tag, api in (("ref", api_ref), ("new", api_new)):
        vv, aa = v.detach().clone().requires_grad_(True), attr.detach().clone().requires_grad_(True)
        index = api.rasterize(vv, vi, H, W)
        depth, bary = api.render(vv, vi, index)
        img = api.interpolate(aa, vi, index, bary)
        out = api.edge_grad_estimator(vv, vi, bary, img, index)
        out.backward(gradient=w)
        res[tag] = dict(index=index, depth=depth.detach(), bary=bary.detach(), img=img.detach(), grad_v=vv.grad, grad_attr=aa.grad)
        del out, img, bary, depth
    r, n = res["ref"], res["new"]
    rep = {"against": "reference CUDA kernels (oracle/_ref), same tensors", "index_equal": bool(th.equal(r["index"], n["index"])),
           "index_mismatches": int((r["index"] != n["index"]).sum()), "err_over_scale": {}, "tolerance": {}}
    ok = rep["index_equal"]
    for k, tol in (("depth", 1e-5), ("bary", 1e-5), ("img", 1e-5), ("grad_v", 5e-5), ("grad_attr", 5e-5)):
        a, e = n[k], r[k]
        scale = float(e.abs().max())
        within = (a - e).abs() <= tol * e.abs() + tol * scale  # False for NaN / Inf
        finite = bool(th.isfinite(a).all())
        rep["err_over_scale"][k] = float(((a - e).abs().max() / max(scale, 1e-30))) if finite else float("nan")
        rep["tolerance"][k] = tol
        ok = ok and finite and bool(within.all())
    rep["ok"] = bool(ok)
    return rep


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="new", choices=["new", "reference"])
    ap.add_argument("--config", type=int, default=4, choices=[2, 3, 4, 5])
    ap.add_argument("--overdraw", action="store_true", help="two sheets, the second rotated 7 degrees (occlusion + intersections)")
    ap.add_argument("--regions", type=int, default=5, help="how many times the K-step region is timed (median reported)")
    ap.add_argument("--transport", default="auto", choices=["auto", "nccl", "multimem"])
    ap.add_argument("--bg-ctas", type=int, default=None,
                    help="multimem transport: CTAs of the exchanges that overlap the backward (default: a quarter of the SMs; "
                         "0 = full wave for every exchange)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the informational legs (cuda graph, with_loss, host paths)")
    ap.add_argument("--dispatch", default="auto", choices=["auto", "torch", "ctypes"],
                    help="host path of the public API: dispatcher ops (csrc/torch_shim.cpp) or ctypes; auto = the package default")
    args = ap.parse_args()

    # stdout carries exactly ONE line (the JSON): everything else that libraries print there (e.g. NCCL's
    # version banner) is routed to stderr at the file-descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg, overdraw = args.config, args.overdraw
    c, C = config_of(cfg)
    F_total = c["F"] * (2 if overdraw else 1)
    workload = (f"BASELINE config {cfg}{' overdraw-2' if overdraw else ''}: {F_total} triangles, {c['W']}x{c['H']}, "
                f"batch {c['N']} per GPU, {C} vertex attributes + edge_grad, synthetic "
                f"{'two-triangle demo scene' if cfg == 2 else 'jittered grid mesh'}")
    metric = METRIC if (cfg == 4 and not overdraw) else METRIC_OTHER.format(cfg=cfg, od=" overdraw-2" if overdraw else "")

    if args.impl == "reference":
        if rank != 0:
            return 0
        r = run_reference_cpu(cfg, max(args.steps, 1), args.warmup, overdraw=overdraw)
        line = {"metric": metric, "value": r["value"], "unit": "Mpix/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
                "config": {"workload": workload, "note": "reference CPU kernels on the host cores, bounded sample: "
                           + r["sample"] + " (Mpix/s is a rate, so the sample compares with the whole-step arm)"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "transform_only")},
                "e2e": {"value": r["value"], "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        emit(line)
        return 0

    assert th.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    th.cuda.set_device(local_rank)
    dev = th.device("cuda", local_rank)
    import drtk_b200
    from drtk_b200 import _ops, torch_ops
    from drtk_b200 import dist as ddist
    import torch.distributed as dist
    if args.dispatch != "auto":
        torch_ops.set_mode(args.dispatch)
    host_path = "dispatcher ops (torch_shim, C++ autograd)" if torch_ops.enabled() else "ctypes (Python autograd)"
    numa_cpus = ddist.bind_to_gpu_numa_node(local_rank)  # before the pinned buffers are allocated
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- inputs: pinned host copies (e2e) and device-resident copies (value) ----
    v_h, vi_h, attr_h, _, _ = make_inputs(cfg, seed_offset=100 * rank, overdraw=overdraw)
    v_h, attr_h = v_h.pin_memory(), attr_h.pin_memory()
    H, W, N = c["H"], c["W"], c["N"]
    npx_rank = N * H * W
    v_d = v_h.to(dev).requires_grad_(True)
    attr_d = attr_h.to(dev).requires_grad_(True)
    vi_d = vi_h.to(dev)  # topology: resident on the device, not part of a step's input
    w = th.rand((N, C, H, W), device=dev, generator=th.Generator(device=dev).manual_seed(1000 * cfg + 2))

    # per-op CUDA-event timing hooks (events on the launching stream, inside the timed region)
    op_events = {k: [] for k in KERNELS}
    orig = {}

    def wrap(name, key):
        fn = getattr(_ops, name)
        orig[name] = fn

        def timed(*a, **kw):
            e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **kw)
            e1.record()
            if timing_on[0]:
                op_events[key].append((e0, e1))
            return out
        setattr(_ops, name, timed)

    timing_on = [False]
    for name, key in (("rasterize", "rasterize"), ("render_forward", "render_fwd"), ("render_backward", "render_bwd"),
                      ("interpolate_forward", "interpolate_fwd"), ("interpolate_backward", "interpolate_bwd"),
                      ("edge_grad_backward_fused", "edge_grad_bwd_fused")):
        wrap(name, key)

    # shared-parameter gradients: batch-summed (and all-reduced over ranks) per parameter, from post-accumulate-grad hooks
    reducer = ddist.SharedGradReducer([v_d, attr_d], transport=args.transport, background_ctas=args.bg_ctas)
    exchange = {"on": True}

    def step_device():
        v_d.grad = None
        attr_d.grad = None
        pipeline(drtk_b200, v_d, vi_d, attr_d, w, H, W)
        return reducer.finish()

    main_s = th.cuda.current_stream(dev)
    bucket_h = th.empty((reducer.total,), dtype=th.float32).pin_memory()
    h2d = v_h.numel() * 4 + attr_h.numel() * 4
    d2h = bucket_h.numel() * 4

    # End-to-end step through the public API with HOST buffers.  The copies are part of every step and inside
    # the timed region, pipelined the way a training loop's prefetcher does it: step k+1's inputs (v_pix, attr)
    # are copied pinned host -> device on a copy stream while step k computes, and step k's result (the reduced
    # [V,3] + [V,C] gradients) is read back on a second copy stream while step k+1 runs.  The device inputs are
    # double-buffered; each buffer is a leaf with its own reducer hooks.
    s_in, s_out = th.cuda.Stream(dev), th.cuda.Stream(dev)
    ebuf = []
    for _ in range(2):
        bv = th.empty_like(v_h, device=dev).requires_grad_(True)
        ba = th.empty_like(attr_h, device=dev).requires_grad_(True)
        ebuf.append(dict(v=bv, a=ba, ready=th.cuda.Event(), free=th.cuda.Event(), out_done=th.cuda.Event(),
                         red=ddist.SharedGradReducer([bv, ba], transport=args.transport, background_ctas=args.bg_ctas)))
    estate = {"k": 0, "primed": False}

    def e2e_prefetch(k):
        b = ebuf[k % 2]
        with th.cuda.stream(s_in), th.no_grad():
            s_in.wait_event(b["free"])
            b["v"].copy_(v_h, non_blocking=True)
            b["a"].copy_(attr_h, non_blocking=True)
            b["ready"].record(s_in)

    def step_e2e():
        k = estate["k"]
        if not estate["primed"]:
            for b in ebuf:
                b["free"].record(main_s)
            e2e_prefetch(k)
            estate["primed"] = True
        b = ebuf[k % 2]
        main_s.wait_event(b["ready"])
        main_s.wait_event(b["out_done"])  # this buffer's bucket was read out two steps ago
        e2e_prefetch(k + 1)
        b["v"].grad = None
        b["a"].grad = None
        pipeline(drtk_b200, b["v"], vi_d, b["a"], w, H, W)
        b["red"].finish()
        b["free"].record(main_s)
        s_out.wait_stream(main_s)
        with th.cuda.stream(s_out):
            bucket_h.copy_(b["red"].bucket, non_blocking=True)
            b["out_done"].record(s_out)
        estate["k"] = k + 1

    def e2e_finish():  # the last step's read-back (and the one prefetch in flight) end inside the timed region
        main_s.wait_stream(s_out)
        main_s.wait_stream(s_in)

    def barrier():
        if world > 1:
            dist.barrier()
        th.cuda.synchronize()

    def timed_region(step_fn, steps, finish=None):
        barrier()
        e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            step_fn()
        if finish:
            finish()
        e1.record()
        barrier()
        if sampler is not None:
            sampler.window(w0, time.perf_counter())
        ms = th.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    R = max(args.regions, 1)
    # no collector pauses on the issuing thread inside the timed regions (one region in ~20 showed a 35-ms host hiccup;
    # the median over R regions is the second guard)
    import gc
    gc.collect()
    gc.disable()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step_device()
    timed_region(step_device, args.steps)  # one more untimed pass of K steps: allocator pools and clocks settle
    dev_runs = [timed_region(step_device, args.steps) for _ in range(R)]
    ms_step = statistics.median(dev_runs)
    # per-op CUDA events (roofline, per_op): one more region of the same K steps over the ctypes host path, whose Python
    # launchers can be bracketed op by op (the dispatcher path runs its backward ops inside the C++ autograd engine);
    # same kernels, same stream, same tensors
    was_torch = torch_ops.enabled()
    torch_ops.set_mode("ctypes")
    step_device()
    timing_on[0] = True
    ms_step_per_op_region = timed_region(step_device, args.steps)
    timing_on[0] = False
    torch_ops.set_mode("torch" if was_torch else "ctypes")

    for _ in range(3):
        step_e2e()
    e2e_finish()
    e2e_runs = [timed_region(step_e2e, args.steps, e2e_finish) for _ in range(R)]
    ms_e2e = statistics.median(e2e_runs)
    clocks = sampler.stop() if sampler else None
    gc.enable()

    # the exchange alone (batch sums + all-reduce of both parameters on gradients already in place)
    def exchange_only():
        reducer._make_hook(0)(v_d)
        reducer._make_hook(1)(attr_d)
        reducer.finish()
    for _ in range(3):
        exchange_only()
    ms_exchange = timed_region(exchange_only, 20)
    for b in ebuf:
        b["red"].close()

    # host <-> device copy bandwidth of this box (context for e2e)
    pcie = None
    if rank == 0:
        big_h = th.empty(64 << 20, dtype=th.uint8).pin_memory()
        big_d = th.empty(64 << 20, dtype=th.uint8, device=dev)
        def _bw(fn):
            fn(); th.cuda.synchronize()
            a0, a1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
            a0.record(); fn(); fn(); a1.record(); th.cuda.synchronize()
            return round(2 * (64 << 20) / (a0.elapsed_time(a1) * 1e-3) / 1e9, 1)
        pcie = {"h2d_GBps": _bw(lambda: big_d.copy_(big_h, non_blocking=True)),
                "d2h_GBps": _bw(lambda: big_h.copy_(big_d, non_blocking=True))}

    total_px = npx_rank * world
    value = total_px / (ms_step * 1e-3) / 1e6
    e2e_value = total_px / (ms_e2e * 1e-3) / 1e6

    # ---- per-op breakdown and the roofline of the dominant kernel (rank 0) ----
    per_op = {}
    for k, evs in op_events.items():
        if evs:
            per_op[k] = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    peak, peak_src = load_peaks()
    roofline, breakdown = None, {}
    if per_op:
        for k, ms in per_op.items():
            gb = ALGO_BYTES_PER_PX[k](C) * npx_rank / 1e9
            breakdown[k] = {"ms": round(ms, 4), "algo_GB": round(gb, 4), "GBps": round(gb / (ms * 1e-3), 1),
                            "frac_of_peak": round(gb / (ms * 1e-3) / peak, 4)}
        dom = max(per_op, key=per_op.get)
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and cfg == 4 and not overdraw:
            with open(tp) as f:
                traffic = json.load(f).get(dom)  # ncu dram bytes per launch of the dominant kernel (config 4 capture)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": breakdown[dom]["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": breakdown[dom]["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": ALGO_BYTES_PER_PX[dom](C) * npx_rank}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    line = {
        "metric": metric, "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "new",
        "ms_per_step_regions": [round(x, 4) for x in dev_runs], "ms_per_step_min": min(dev_runs),
        "host_path": host_path, "per_op_region_ms_per_step": round(ms_step_per_op_region, 4),
        "config": {"workload": workload, "global_batch": N * world,
                   "parallelism": f"dp{world} (batch sharded; shared-parameter gradients batch-summed"
                                  + (f" and all-reduced, transport {reducer.transport})" if world > 1 else ")"),
                   "l2": "per-step working set >> 126 MB L2 at configs 3-5 (inputs larger than L2, no explicit flush)",
                   "loss": "none materialised: backward seeded with cotangent w (= gradient of the linear loss (img*w).sum()); see with_loss",
                   "regions": f"{R} timed regions of {args.steps} steps; ms_per_step / value = median region"},
        "e2e": {"value": e2e_value, "unit": "Mpix/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step_regions": [round(x, 4) for x in e2e_runs],
                "note": "pinned host v_pix + attr copied in, the step's result (gradients summed over the batch"
                        + (" and ranks" if world > 1 else "") + ": [V,3] + [V,C]) copied out, every step, inside the timed "
                        "region; copies double-buffered on side streams; topology (vi) stays resident; median of the regions"},
        "gpu_launches": (sum(KERNELS.values()) + REDUCE_KERNELS) * args.steps,
        "clocks": clocks, "roofline": roofline, "per_op": breakdown, "pcie": pcie,
        "shared_grad_exchange": {"transport": reducer.transport, "background_ctas": args.bg_ctas, "ms_isolated": round(ms_exchange, 4),
                                 "bytes": reducer.total * 4, "fallback_reason": getattr(reducer, "fallback_reason", None)},
        "numa_cpus": (f"{numa_cpus[0]}-{numa_cpus[-1]} ({len(numa_cpus)})" if numa_cpus else None),
        "pipeline_algorithmic_GB_per_step": round(sum(ALGO_BYTES_PER_PX[k](C) for k in KERNELS) * npx_rank / 1e9, 3),
    }

    # ---- reference CUDA kernels on the same tensors (context for the >=4x target; not the arm) + parity of the timed tensors ----
    R_api = None
    if world == 1 and not args.no_ref_cuda:
        try:
            from oracle import ref as R_api_mod
            if R_api_mod.available():
                R_api = R_api_mod

                def step_ref():
                    v_d.grad = None; attr_d.grad = None
                    pipeline(R_api, v_d, vi_d, attr_d, w, H, W)
                for _ in range(3):
                    step_ref()
                ref_runs = [timed_region(step_ref, args.steps) for _ in range(R)]
                ms_ref = statistics.median(ref_runs)
                line["reference_cuda"] = {"value": total_px / (ms_ref * 1e-3) / 1e6, "unit": "Mpix/s", "ms_per_step": ms_ref,
                                          "ms_per_step_regions": [round(x, 4) for x in ref_runs],
                                          "note": "unmodified reference CUDA kernels (oracle/_ref, sm_100 build), same tensors, same cotangent, no batch sum"}
                line["parity_check"] = compare_with_reference(drtk_b200, R_api, v_d, vi_d, attr_d, w, H, W)
        except Exception as ex:  # noqa: BLE001
            line["reference_cuda"] = {"unavailable": repr(ex)[:200]}
    v_d.grad = None
    attr_d.grad = None
    reducer.close()

    if world == 1 and not args.no_extras:
        for name, fn in orig.items():  # drop the per-op event hooks
            setattr(_ops, name, fn)
        # ---- the same step captured once in a CUDA graph and replayed (informational: launch + Python overhead out) ----
        try:
            gv_s = v_d.detach().clone().requires_grad_(True)
            ga_s = attr_d.detach().clone().requires_grad_(True)

            def graph_step():
                index = drtk_b200.rasterize(gv_s, vi_d, H, W)
                _, bary = drtk_b200.render(gv_s, vi_d, index)
                img = drtk_b200.interpolate(ga_s, vi_d, index, bary)
                img = drtk_b200.edge_grad_estimator(gv_s, vi_d, bary, img, index)
                gvv, gaa = th.autograd.grad(img, (gv_s, ga_s), grad_outputs=w)
                return ddist.batch_sum(gvv), ddist.batch_sum(gaa)

            side = th.cuda.Stream(dev)
            side.wait_stream(main_s)
            with th.cuda.stream(side):
                for _ in range(2):
                    graph_step()
            main_s.wait_stream(side)
            graph = th.cuda.CUDAGraph()
            with th.cuda.graph(graph):
                graph_out = graph_step()
            for _ in range(2):
                graph.replay()
            ms_graph = timed_region(graph.replay, args.steps)
            line["cuda_graph"] = {"ms_per_step": ms_graph, "value": total_px / (ms_graph * 1e-3) / 1e6, "unit": "Mpix/s",
                                  "note": "one forward+backward step captured with torch.cuda.graph and replayed; same kernels"}
            del graph, graph_out
        except Exception as ex:  # noqa: BLE001
            line["cuda_graph"] = {"unavailable": repr(ex)[:200]}

        # ---- SURVEY 8(d)'s literal pipeline: loss = (img * w).sum(); loss.backward()  (torch's own kernels on the big tensors) ----
        try:
            def loss_step(api):
                v_d.grad = None; attr_d.grad = None
                index = api.rasterize(v_d, vi_d, H, W)
                _, bary = api.render(v_d, vi_d, index)
                img = api.interpolate(attr_d, vi_d, index, bary)
                img = api.edge_grad_estimator(v_d, vi_d, bary, img, index)
                (img * w).sum().backward()
            for _ in range(2):
                loss_step(drtk_b200)
            ms_l = timed_region(lambda: loss_step(drtk_b200), args.steps)
            wl = {"ms_per_step": ms_l, "value": total_px / (ms_l * 1e-3) / 1e6, "unit": "Mpix/s",
                  "note": "informational: the loss materialised with torch ops ((img*w).sum().backward()) inside the step, both arms"}
            if R_api is not None:
                for _ in range(2):
                    loss_step(R_api)
                ms_lr = timed_region(lambda: loss_step(R_api), args.steps)
                wl["reference_cuda_ms_per_step"] = ms_lr
            line["with_loss"] = wl
        except Exception as ex:  # noqa: BLE001
            line["with_loss"] = {"unavailable": repr(ex)[:200]}

        # ---- host paths: wall-clock per step of the Python host (ctypes) vs the dispatcher ops (torch_shim) vs the reference ----
        try:
            from drtk_b200 import torch_ops

            class _Shim:  # the public functions, forced through torch.ops.drtk_b200_*_ext
                @staticmethod
                def rasterize(v, vi_, h, w_):
                    return torch_ops.rasterize(v, vi_[None].expand(v.shape[0], -1, -1) if vi_.ndim == 2 else vi_, h, w_)[1]

                @staticmethod
                def render(v, vi_, index):
                    d, b = torch_ops.render(v, vi_[None].expand(v.shape[0], -1, -1) if vi_.ndim == 2 else vi_, index)
                    return d, b

                @staticmethod
                def interpolate(a, vi_, index, bary):
                    return torch_ops.interpolate(a, vi_[None].expand(a.shape[0], -1, -1) if vi_.ndim == 2 else vi_, index, bary)

                @staticmethod
                def edge_grad_estimator(v, vi_, bary, img, index):
                    return torch_ops.edge_grad_estimator_fused(v, vi_[None].expand(v.shape[0], -1, -1) if vi_.ndim == 2 else vi_,
                                                               bary.detach(), img, index)

            def wall(api, iters):
                def one():
                    v_d.grad = None; attr_d.grad = None
                    pipeline(api, v_d, vi_d, attr_d, w, H, W)
                for _ in range(3):
                    one()
                th.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(iters):
                    one()
                t_issue = time.perf_counter() - t0
                th.cuda.synchronize()
                return {"issue_us_per_step": round(t_issue / iters * 1e6, 1), "wall_us_per_step": round((time.perf_counter() - t0) / iters * 1e6, 1)}
            iters = 50 if cfg <= 3 else 10
            torch_ops.set_mode("ctypes")
            hp = {"ctypes_python_autograd": wall(drtk_b200, iters)}
            torch_ops.set_mode("torch" if was_torch else "ctypes")
            if torch_ops.available():
                hp["dispatcher_ops_cpp_autograd"] = wall(_Shim, iters)
            if R_api is not None:
                hp["reference_cuda"] = wall(R_api, iters)
            hp["note"] = "issue = host time to enqueue one forward+backward step; wall = including the device work"
            line["host_paths"] = hp
        except Exception as ex:  # noqa: BLE001
            line["host_paths"] = {"unavailable": repr(ex)[:200]}

    if world == 1 and not args.no_cpu_baseline:
        r = run_reference_cpu(cfg, steps=3, warmup=1, overdraw=overdraw)
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "transform_only")}
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
