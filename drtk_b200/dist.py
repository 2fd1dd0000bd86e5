"""Multi-GPU plumbing for the rasterisation path: one process per GPU, batch sharded.

The reference has no distributed layer.  Every kernel of the path treats batch
items independently (e.g. `n = index / (H*W)` decode, src/render/render_kernel.cu:58-60), so the
batch dimension shards across ranks with NO data-path collective.  The only exchange is the
gradient of parameters SHARED by all batch items (a common mesh / attribute table): each rank
sums its local batch and the ranks all-reduce the [V,3] (+[V,C]) result.  Per-item parameters
need no collective at all.

Two transports for that exchange:
  "nccl"      our batch-sum kernel writes the local sums into one flat bucket, `dist.all_reduce` on it;
  "multimem"  ONE kernel does the local batch sum and the sum over ranks: `drtk_b200_batch_sum_allreduce` pushes
              the local sums into every rank's copy of a symmetric-memory bucket with `multimem.red.add.f32`
              through the NVSwitch multicast address (flag barriers over peer memory, no NCCL call).
`SharedGradReducer` issues the exchange of each parameter from a post-accumulate-grad hook on a side stream, so
the attribute-table gradients travel while render's and edge_grad's backward kernels still run.
"""
import ctypes
import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch as th
import torch.distributed as dist


def shard_batch(n_global: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[begin, end) of the contiguous slice of the batch owned by `rank` (sizes differ by <= 1)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(n_global, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def _world(group=None) -> int:
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def batch_sum(x: th.Tensor, out: Optional[th.Tensor] = None) -> th.Tensor:
    """sum over dim 0 of a per-item gradient [N, ...] -> [...].  CUDA float32: the library's one-launch kernel
    (128-bit accesses, written into `out` -- typically a slice of a communication bucket); anything else: torch."""
    if out is None:
        out = th.empty(x.shape[1:], dtype=x.dtype, device=x.device)
    if x.is_cuda and x.dtype == th.float32 and x.dim() >= 1 and x[0].is_contiguous() and out.is_contiguous():
        from . import _lib
        lib = _lib.load()
        N = x.shape[0]
        M = out.numel()
        args = (_lib.ptr(x), N, M, x.stride(0) if N > 1 else M, _lib.ptr(out), th.cuda.current_stream(x.device).cuda_stream)
        if th.cuda.current_device() == x.device.index:  # the common case: no device-guard round trip on the host path
            rc = lib.drtk_b200_batch_sum(*args)
        else:
            with th.cuda.device(x.device):
                rc = lib.drtk_b200_batch_sum(*args)
        _lib.check(rc, "batch_sum()")
    else:
        th.sum(x, dim=0, out=out)
    return out


def allreduce_shared_grads(grads: Iterable[Optional[th.Tensor]], group=None, async_op: bool = False):
    """Sum per-item gradients [N_local, ...] over the local batch, then all-reduce (SUM) across
    ranks.  All tensors travel in ONE flat fp32 bucket (one collective launch: the payload is a few MB,
    latency bound on NVLink 5 / NVSwitch).  Returns the list of reduced [...] tensors (views into
    the bucket) and, with async_op=True, the work handle to wait on."""
    gl: List[th.Tensor] = [g for g in grads if g is not None]
    if not gl:
        return [], None
    sizes = [g[0].numel() for g in gl]
    flat = th.empty((sum(sizes),), dtype=gl[0].dtype, device=gl[0].device)
    outs, off = [], 0
    for g, m in zip(gl, sizes):
        o = flat[off:off + m].view(g.shape[1:])
        batch_sum(g, o)
        outs.append(o)
        off += m
    work = None
    if _world(group) > 1:
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
    return outs, work


class _MultimemBucket:
    """A symmetric-memory bucket of 2 x `numel` floats (two halves used by alternate backward passes) plus the flag
    array of the in-kernel rank barrier."""

    def __init__(self, numel: int, device: th.device, group):
        import torch.distributed._symmetric_memory as symm_mem
        from . import _lib
        self.lib = _lib.load()
        self._lib = _lib
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        self.numel = numel
        grid = int(self.lib.drtk_b200_batch_sum_allreduce_grid())
        self.grid = grid
        self.bucket2 = symm_mem.empty(2 * numel, dtype=th.float32, device=device)
        self.bucket2.zero_()
        self.flags = symm_mem.empty(self.world * grid, dtype=th.int32, device=device)
        self.flags.zero_()
        self.h_bucket = symm_mem.rendezvous(self.bucket2, self.group)
        self.h_flags = symm_mem.rendezvous(self.flags, self.group)
        if not getattr(self.h_bucket, "multicast_ptr", 0):
            raise RuntimeError("symmetric memory has no multicast address on this system (NVLS unavailable)")
        self.peer_flags = (ctypes.c_void_p * self.world)(*[int(p) for p in self.h_flags.buffer_ptrs])
        self.timeout = th.zeros((1,), dtype=th.int32, device=device)
        self.epoch = 0
        self.half = 0
        th.cuda.synchronize(device)
        dist.barrier(self.group)  # every rank's flags and buckets are zero before anybody touches a peer's

    @property
    def bucket(self) -> th.Tensor:
        """The half that holds (or is receiving) the current pass's sums."""
        return self.bucket2[self.half * self.numel:(self.half + 1) * self.numel]

    def reduce(self, x: th.Tensor, offset: int, numel: int, stream: th.cuda.Stream, max_ctas: int = 0) -> None:
        """current half[offset : offset+numel] = sum over ranks of sum_n x[n]; the same range of the other half of
        THIS rank is zero-filled for the next pass (one kernel, one cross-rank barrier).  max_ctas > 0 caps the grid
        (an exchange that overlaps other kernels); every rank must pass the same value."""
        N = x.shape[0]
        acc = self.half * self.numel + offset
        zero = (1 - self.half) * self.numel + offset
        rc = self.lib.drtk_b200_batch_sum_allreduce(
            self._lib.ptr(x), N, numel, x.stride(0) if N > 1 else numel, self.bucket2.data_ptr() + 4 * zero,
            int(self.h_bucket.multicast_ptr) + 4 * acc, self.peer_flags, self.rank, self.world, self.epoch,
            self.timeout.data_ptr(), int(max_ctas), stream.cuda_stream)
        self._lib.check(rc, "batch_sum_allreduce()")
        self.epoch = (self.epoch + 1) & 0xFFFFFFFF

    def next_pass(self) -> None:
        self.half ^= 1

    def check(self) -> None:
        if int(self.timeout.item()) != 0:
            raise RuntimeError("drtk_b200: a rank did not arrive at the in-kernel all-reduce barrier (timeout)")


class SharedGradReducer:
    """All-reduce the batch-summed gradient of each SHARED parameter as soon as autograd has finished it, instead of
    after the whole backward: `vert_attributes.grad` is complete when interpolate's backward returns, while render's
    and edge_grad's backward kernels (the gradient of `v_pix`) are still to run, so its exchange hides behind them.

        reducer = SharedGradReducer([v_pix, attr])                 # leaves of shape [N_local, ...]
        loss.backward()                                            # hooks fire per parameter
        grad_v, grad_attr = reducer.finish()                       # [...] tensors, summed over batch and ranks

    transport: "nccl" | "multimem" | "auto" (multimem when the process group is NCCL on CUDA and symmetric memory offers
    a multicast address, else nccl; "auto" never raises).  On CUDA the work is issued on a side stream (ordered after
    the gradient's producer through an event; the caller's stream only waits in `finish()`).  Without an initialised
    process group it degenerates to the local batch sums.  The returned tensors are views into the reducer's bucket:
    consume (or copy) them before the next backward pass, and call `finish()` once per backward pass (the multimem
    transport alternates between two bucket halves per pass)."""

    def __init__(self, params: Sequence[th.Tensor], group=None, transport: str = "auto",
                 background_ctas: Optional[int] = None):
        self.params = list(params)
        self.group = group
        # multimem transport: the exchange of the parameter whose gradient arrives LAST in a backward pass is on the
        # critical path and runs on a full wave of CTAs; the earlier ones overlap the rest of the backward and run on
        # `background_ctas` CTAs (default: a quarter of the SMs), because their CTAs sit at the rank barrier until the
        # slowest rank arrives and would otherwise take registers and issue slots from the kernels they overlap.
        # The order is learned from the previous pass (identical on every rank: same autograd graph).
        self.background_ctas = background_ctas
        self._fired: List[int] = []
        self._last_fired: Optional[int] = None
        self.sizes = [p[0].numel() for p in self.params]
        self.offsets = [sum(self.sizes[:i]) for i in range(len(self.sizes))]
        # segments start on 16-byte boundaries so that every one takes the 128-bit path
        self.offsets = []
        off = 0
        for m in self.sizes:
            self.offsets.append(off)
            off += (m + 3) // 4 * 4
        self.total = off
        dev = self.params[0].device
        self.cuda = dev.type == "cuda"
        self.mm = None
        self.transport = "local"
        if self._distributed():
            self.transport = "nccl"
            if transport not in ("auto", "nccl", "multimem"):
                raise ValueError(f"transport {transport!r}")
            if transport in ("auto", "multimem") and self.cuda and all(p.dtype == th.float32 for p in self.params):
                try:
                    self.mm = _MultimemBucket(self.total, dev, group)
                    self.transport = "multimem"
                except Exception as ex:  # noqa: BLE001
                    if transport == "multimem":
                        raise
                    self.fallback_reason = repr(ex)[:200]
        self._bucket = None if self.mm is not None else th.zeros((self.total,), dtype=self.params[0].dtype, device=dev)
        self._last = None  # the bucket (half) that holds the results of the last finished pass
        self._pending = {}
        # the side stream only pays when there is a cross-rank exchange to hide; without one its three extra stream
        # calls per parameter are pure host overhead (they dominate the step of a 512^2 demo scene)
        self._overlap = self.cuda and self._distributed()
        self._side = th.cuda.Stream(dev) if self._overlap else None
        self._handles = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _distributed(self) -> bool:
        return _world(self.group) > 1

    @property
    def bucket(self) -> th.Tensor:
        """Flat bucket holding the results of the last finished pass (before the first `finish()`: the current one)."""
        if self._last is not None:
            return self._last
        return self.mm.bucket if self.mm is not None else self._bucket

    def _current(self) -> th.Tensor:
        return self.mm.bucket if self.mm is not None else self._bucket

    def _segment(self, i: int, flat: Optional[th.Tensor] = None) -> th.Tensor:
        flat = self._current() if flat is None else flat
        return flat[self.offsets[i]:self.offsets[i] + self.sizes[i]].view(self.params[i].shape[1:])

    def _make_hook(self, i):
        def hook(p):
            g = p.grad
            if self.cuda and not self._overlap:  # single process: the local sums run in stream order, no side stream
                work = self._exchange(i, g)
            elif self.cuda:
                self._side.wait_stream(th.cuda.current_stream(p.device))
                with th.cuda.stream(self._side):
                    work = self._exchange(i, g)
                g.record_stream(self._side)
            else:
                work = self._exchange(i, g)
            self._pending[i] = work
        return hook

    def _exchange(self, i: int, g: th.Tensor):
        seg = self._segment(i)
        if self.mm is not None:
            self._fired.append(i)
            bg = 0
            if self._last_fired is not None and i != self._last_fired:
                bg = self.background_ctas if self.background_ctas is not None else max(1, self.mm.grid // 4)
            self.mm.reduce(g, self.offsets[i], self.sizes[i], th.cuda.current_stream(g.device), max_ctas=bg)
            return None
        batch_sum(g, seg)
        if self._distributed():
            return dist.all_reduce(seg, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        return None

    def finish(self) -> List[Optional[th.Tensor]]:
        """Wait for the exchanges of this backward pass; returns the reduced gradients in parameter order (None for a
        parameter that received no gradient)."""
        out: List[Optional[th.Tensor]] = []
        flat = self._current()
        for i in range(len(self.params)):
            if i not in self._pending:
                out.append(None)
                continue
            work = self._pending.pop(i)
            if work is not None:
                work.wait()  # on CUDA: the current stream waits for the collective, the host does not block
            out.append(self._segment(i, flat))
        if self._overlap:
            th.cuda.current_stream(self.params[0].device).wait_stream(self._side)
        self._last = flat
        if self.mm is not None:
            self.mm.next_pass()
            if self._fired:
                self._last_fired = self._fired[-1]
            self._fired = []
        return out

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


# round-1 name
OverlappedSharedGradReducer = SharedGradReducer


def bind_to_gpu_numa_node(device_index: int) -> Optional[List[int]]:
    """Pin this process (and the pinned host buffers it allocates afterwards) to the CPUs NVML reports as local to the
    GPU: with one process per GPU the host <-> device copies of every rank then stay on their own memory controller
    and PCIe root instead of crowding NUMA node 0.  Returns the CPU list, or None when NVML / affinity is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            try:
                phys = int(vis.split(",")[device_index])
            except Exception:  # noqa: BLE001
                phys = device_index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        ncpu = os.cpu_count() or 1
        words = (ncpu + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [w * 64 + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1 and w * 64 + b < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:  # noqa: BLE001
        return None
    return None
