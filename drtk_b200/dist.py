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
