"""CPU suite: the N>1 path (batch sharding + shared-gradient all-reduce) with world_size 2 on gloo."""
import os
import socket

import torch as th
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, V, C, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from drtk_b200 import dist as ddist
    gen = th.Generator().manual_seed(1234)
    gv_all = th.randn(n_global, V, 3, generator=gen)      # per-item gradients of the whole job
    ga_all = th.randn(n_global, V, C, generator=gen)
    b, e = ddist.shard_batch(n_global, rank, world)
    (gv, ga), work = ddist.allreduce_shared_grads([gv_all[b:e], ga_all[b:e]], async_op=True)
    work.wait()
    th.save((gv, ga), os.path.join(out_dir, f"r{rank}.pt"))
    # the overlapped reducer: hooks fire as autograd finishes each parameter's gradient
    pv = th.zeros(e - b, V, 3, requires_grad=True)
    pa = th.zeros(e - b, V, C, requires_grad=True)
    red = ddist.OverlappedSharedGradReducer([pv, pa])
    for _ in range(2):  # two backward passes: the reducer is reusable
        pv.grad = None; pa.grad = None
        ((pv * gv_all[b:e]).sum() + (pa * ga_all[b:e]).sum() + (pv * gv_all[b:e]).sum()).backward()  # v gets two contributions
        ov, oa = red.finish()
    red.close()
    th.save((ov, oa), os.path.join(out_dir, f"o{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_shared_gradient_allreduce(tmp_path):
    world, n_global, V, C = 2, 5, 17, 4
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_global, V, C, str(tmp_path)), nprocs=world, join=True)
    gen = th.Generator().manual_seed(1234)
    gv_all = th.randn(n_global, V, 3, generator=gen)
    ga_all = th.randn(n_global, V, C, generator=gen)
    for r in range(world):
        gv, ga = th.load(os.path.join(str(tmp_path), f"r{r}.pt"))
        assert th.allclose(gv, gv_all.sum(0), atol=1e-5)
        assert th.allclose(ga, ga_all.sum(0), atol=1e-5)
        ov, oa = th.load(os.path.join(str(tmp_path), f"o{r}.pt"))
        assert th.allclose(ov, 2 * gv_all.sum(0), atol=1e-5) and th.allclose(oa, ga_all.sum(0), atol=1e-5)


def test_overlapped_reducer_without_process_group():
    from drtk_b200 import dist as ddist
    p = th.zeros(3, 5, 2, requires_grad=True)
    w = th.arange(30.0).view(3, 5, 2)
    red = ddist.OverlappedSharedGradReducer([p])
    (p * w).sum().backward()
    (g,) = red.finish()
    red.close()
    assert th.equal(g, w.sum(0))
    assert red.finish() == [None]


def test_reducer_background_grid_follows_the_firing_order(monkeypatch):
    """multimem transport, host logic only (a stub stands in for the symmetric-memory bucket): the exchange of the parameter
    whose gradient arrived LAST in the previous pass runs on a full wave (max_ctas 0), the earlier ones on the background
    grid; the first pass, whose order is unknown, runs everything on a full wave.  Every rank derives the same values from
    the same autograd graph -- a mismatch would leave the ranks' in-kernel barriers with different grids."""
    import torch as th
    from drtk_b200 import dist as ddist

    calls = []

    class StubBucket:
        grid = 148

        def __init__(self, numel):
            self.flat = th.zeros((numel,))

        @property
        def bucket(self):
            return self.flat

        def reduce(self, x, offset, numel, stream, max_ctas=0):
            calls.append((offset, max_ctas))
            self.flat[offset:offset + numel] = x.sum(0).reshape(-1)

        def next_pass(self):
            pass

    monkeypatch.setattr(ddist.th.cuda, "current_stream", lambda device=None: None)
    for background, expect in ((None, 148 // 4), (16, 16), (0, 0)):
        v = th.zeros((2, 5, 3), requires_grad=True)
        a = th.zeros((2, 5, 4), requires_grad=True)
        red = ddist.SharedGradReducer([v, a], background_ctas=background)
        red.mm = StubBucket(red.total)
        for it in range(3):
            calls.clear()
            v.grad = None
            a.grad = None
            (a * 2.0).sum().backward()   # a fires first ...
            (v * 3.0).sum().backward()   # ... v last
            gv, ga = red.finish()
            assert th.allclose(gv, th.full((5, 3), 6.0)) and th.allclose(ga, th.full((5, 4), 4.0))
            off_v, off_a = red.offsets
            if it == 0:
                assert calls == [(off_a, 0), (off_v, 0)]
            else:
                assert calls == [(off_a, expect), (off_v, 0)]
        red.close()
