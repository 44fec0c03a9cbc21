"""Multi-GPU: one process per GPU, bundles sharded by ray index, ONE collective per step.

Rays are independent (no reference code couples two rays), so the path shards with no
data-path exchange: rank r traces rays [lo_r, hi_r) of the bundle against a replicated surface
table (a few KB).  The only exchange step is the reduction of what the trace ACCUMULATES over
rays — sensor images and parameter gradients (and scalar loss terms) — which are packed into one
flat fp32 buffer and summed with a single NCCL all-reduce over NVLink/NVSwitch (gloo on CPU for
the tests).  Per-ray outputs stay on the rank that owns the rays.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the default
    process group when WORLD_SIZE > 1 (nccl with CUDA, gloo otherwise)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: Optional[int] = None, world_size: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous, balanced index range of this rank: sizes differ by at most one ray and the
    ranges tile [0, n) exactly (also for n < world_size, where trailing ranks get an empty shard)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    base, rem = divmod(int(n), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays, rank: Optional[int] = None, world_size: Optional[int] = None):
    lo, hi = shard_bounds(len(rays), rank, world_size)
    return rays[lo:hi]


class FlatReducer:
    """Packs tensors into one flat fp32 buffer, all-reduces it (SUM) once, unpacks in place.

    Usage per step:  ``red = FlatReducer(); red.add(image); red.add(grad_c1); ...; red.reduce()``.
    With world_size 1 ``reduce`` is a no-op (no copy)."""

    def __init__(self, group=None):
        self.group = group
        self.items: List[torch.Tensor] = []

    def add(self, t: Optional[torch.Tensor]):
        if t is not None:
            self.items.append(t)
        return t

    def extend(self, ts: Iterable[Optional[torch.Tensor]]):
        for t in ts:
            self.add(t)

    def reduce(self, async_op: bool = False):
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1 \
                or not self.items:
            return None
        if len(self.items) == 1 and self.items[0].is_contiguous() and self.items[0].dtype == torch.float32:
            return dist.all_reduce(self.items[0], op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)
        flat = torch.cat([t.reshape(-1).to(torch.float32) for t in self.items])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=async_op)

        def unpack():
            off = 0
            for t in self.items:
                n = t.numel()
                t.copy_(flat[off:off + n].view_as(t))
                off += n

        if async_op:
            class _W:
                def wait(self_inner):
                    work.wait()
                    unpack()
            return _W()
        unpack()
        return None


class OverlappedReducer:
    """All-reduce of per-step accumulators (sensor images) on a side stream, so that the collective of step k runs
    under the trace of step k+1 instead of after it on the same stream.

    ``submit(tensors)`` — call right after the kernels that produced them were enqueued on the current stream; the
    side stream waits for that point, then runs ONE flat all-reduce (SUM, in place).  ``drain()`` makes the current
    stream wait for every submitted reduction (call before reading the results / at the end of a timed region).
    Each step must hand over FRESH tensors (the trace ops allocate their images per call), never a buffer a later
    step writes while the reduction may still be in flight."""

    def __init__(self, device, group=None):
        self.group = group
        self.stream = torch.cuda.Stream(device=device)
        self.pending: List[Tuple[torch.cuda.Event, list]] = []

    def submit(self, tensors: Iterable[Optional[torch.Tensor]]):
        items = [t for t in tensors if t is not None]
        if not items or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            return
        cur = torch.cuda.current_stream(self.stream.device)
        ready = torch.cuda.Event()
        ready.record(cur)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            red = FlatReducer(self.group)
            red.extend(items)
            red.reduce()
            done = torch.cuda.Event()
            done.record(self.stream)
        for t in items:
            t.record_stream(self.stream)
        self.pending.append((done, items))
        if len(self.pending) > 4:                       # bound the number of images kept alive
            ev, _ = self.pending.pop(0)
            cur.wait_event(ev)

    def drain(self):
        cur = torch.cuda.current_stream(self.stream.device)
        for ev, _ in self.pending:
            cur.wait_event(ev)
        self.pending = []


def bind_to_gpu_numa(local_rank: int) -> Optional[List[int]]:
    """Pin this process to the CPU cores NVML reports as local to GPU ``local_rank`` (its NUMA node), BEFORE pinned
    host buffers are allocated: first-touch then places the staging memory on that node and the H2D copies of several
    ranks do not all cross one socket's memory controller.  Returns the core list, or None when NVML / affinity is
    unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(local_rank))
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cores = [64 * w + b for w, word in enumerate(mask) for b in range(64) if (int(word) >> b) & 1]
        allowed = sorted(set(cores) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return allowed
    except Exception:
        return None
    return None


def allreduce_scene_results(sensors: Sequence, params: Iterable[torch.nn.Parameter] = (), extra: Sequence = ()):
    """Sum sensor images, parameter gradients and extra accumulators (loss moments) over ranks
    with one collective.  Call after ``loss.backward()`` on every rank."""
    red = FlatReducer()
    for s in sensors:
        red.add(getattr(s, "image", None))
    for p in params:
        if p.grad is not None:
            red.add(p.grad)
    red.extend(extra)
    red.reduce()
