"""Multi-GPU check of drtk_b200.dist.SharedGradReducer (run under torchrun, one process per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py
For each transport (nccl, multimem) the reduced gradients must equal an independent NCCL all-reduce of torch.sum over the
batch, over several backward passes (the multimem bucket is reused: zeroing / epochs / barriers are exercised), and the
exchange is timed in isolation.  Prints one line per transport on rank 0; exit code 1 on a mismatch."""
import os, sys, time
import torch as th
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drtk_b200 import dist as ddist

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
th.cuda.set_device(local)
dev = th.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, V, C = 8, 50625, 16
ok = True
for transport in ("nccl", "multimem"):
    try:
        v = th.zeros((N, V, 3), device=dev, requires_grad=True)
        a = th.zeros((N, V, C), device=dev, requires_grad=True)
        red = ddist.SharedGradReducer([v, a], transport=transport)
    except Exception as ex:  # noqa: BLE001
        if rank == 0:
            print(f"{transport:9s} UNAVAILABLE: {repr(ex)[:300]}", flush=True)
        continue
    worst = 0.0
    for it in range(4):
        g = th.Generator(device=dev).manual_seed(1000 * it + rank)
        gv = th.randn((N, V, 3), device=dev, generator=g)
        ga = th.randn((N, V, C), device=dev, generator=g)
        v.grad = None; a.grad = None
        (v * gv).sum().backward(inputs=[v])      # hooks fire: v first here, then a (order differs from the pipeline)
        (a * ga).sum().backward(inputs=[a])
        rv, ra = red.finish()
        ev, ea = gv.sum(0), ga.sum(0)
        dist.all_reduce(ev); dist.all_reduce(ea)
        th.cuda.synchronize()
        for x, e in ((rv, ev), (ra, ea)):
            err = float((x - e).abs().max() / e.abs().max())
            worst = max(worst, err)
    if red.mm is not None:
        red.mm.check()
    # timing of the exchange alone
    def once():
        red._make_hook(0)(v); red._make_hook(1)(a); red.finish()
    for _ in range(5):
        once()
    th.cuda.synchronize(); dist.barrier()
    e0, e1 = th.cuda.Event(enable_timing=True), th.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        once()
    e1.record(); th.cuda.synchronize()
    good = worst < 1e-5
    ok = ok and good
    if rank == 0:
        print(f"{red.transport:9s} world {world}: max rel err {worst:.2e} {'OK' if good else 'BAD'}; exchange {e0.elapsed_time(e1) / 50 * 1e3:.1f} us "
              f"({(V * 3 + V * C) * 4 / 1e6:.2f} MB)", flush=True)
    red.close()
    del red
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
