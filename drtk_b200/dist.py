"""Multi-GPU plumbing for the rasterisation path: one process per GPU, batch sharded.

The reference has no distributed layer.  Every kernel of the path treats batch
items independently (e.g. `n = index / (H*W)` decode, src/render/render_kernel.cu:58-60), so the
batch dimension shards across ranks with NO data-path collective.  The only exchange is the
gradient of parameters SHARED by all batch items (a common mesh / attribute table): each rank
sums its local batch and the ranks all-reduce the [V,3] (+[V,C]) result.  Per-item parameters
need no collective at all.
"""
from typing import Iterable, List, Optional, Tuple

import torch as th
import torch.distributed as dist


def shard_batch(n_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the contiguous slice of the batch owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_global, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_shared_grads(grads: Iterable[Optional[th.Tensor]], group=None, async_op: bool = False):
    """Sum per-item gradients [N_local, ...] over the local batch, then all-reduce (SUM) across
    ranks.  All tensors travel in ONE flat fp32 bucket (one NCCL launch: the payload is a few MB,
    latency bound on NVLink 5 / NVSwitch).  Returns the list of reduced [...] tensors (views into
    the bucket) and, with async_op=True, the work handle to wait on."""
    gl: List[th.Tensor] = [g for g in grads if g is not None]
    if not gl:
        return [], None
    local = [g.sum(dim=0) for g in gl]
    flat = th.cat([x.reshape(-1) for x in local])
    work = None
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    outs, off = [], 0
    for x in local:
        outs.append(flat[off:off + x.numel()].view_as(x))
        off += x.numel()
    return outs, work


class OverlappedSharedGradReducer:
    """All-reduce the batch-summed gradient of each SHARED parameter as soon as autograd has finished it, instead of
    after the whole backward: `vert_attributes.grad` is complete when interpolate's backward returns, while render's
    and edge_grad's backward kernels (the gradient of `v_pix`) are still to run, so its exchange hides behind them.

        reducer = OverlappedSharedGradReducer([v_pix, attr])       # leaves of shape [N_local, ...]
        loss.backward()                                            # hooks fire per parameter
        grad_v, grad_attr = reducer.finish()                       # [...] tensors, summed over batch and ranks

    On CUDA the reduction is issued on a side stream (the collective orders itself after the gradient's producer
    through an event, the caller's stream only waits in `finish()`).  Works without an initialised process group
    (plain batch sums).  The gloo world-size-2 test covers the logic; the NCCL path has not been timed yet, so
    `bench.py` still uses the single bucketed call of `allreduce_shared_grads`."""

    def __init__(self, params: Iterable[th.Tensor], group=None):
        self.params = list(params)
        self.group = group
        self._pending = {}
        self._side = {}
        self._handles = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _distributed(self) -> bool:
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1

    def _make_hook(self, i):
        def hook(p):
            if p.is_cuda:
                side = self._side.setdefault(p.device, th.cuda.Stream(p.device))
                side.wait_stream(th.cuda.current_stream(p.device))
                with th.cuda.stream(side):
                    local = p.grad.sum(dim=0)
                    work = dist.all_reduce(local, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self._distributed() else None
                p.grad.record_stream(side)
            else:
                local = p.grad.sum(dim=0)
                work = dist.all_reduce(local, op=dist.ReduceOp.SUM, group=self.group, async_op=True) if self._distributed() else None
            self._pending[i] = (local, work)
        return hook

    def finish(self) -> List[Optional[th.Tensor]]:
        """Wait for the exchanges of this backward pass; returns the reduced gradients in parameter order (None for a
        parameter that received no gradient)."""
        out: List[Optional[th.Tensor]] = []
        for i, p in enumerate(self.params):
            local, work = self._pending.pop(i, (None, None))
            if work is not None:
                work.wait()  # on CUDA: the current stream waits for the collective, the host does not block
            if local is not None and local.is_cuda:
                cur = th.cuda.current_stream(local.device)
                cur.wait_stream(self._side[local.device])
                local.record_stream(cur)
            out.append(local)
        return out

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []
