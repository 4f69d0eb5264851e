"""Multi-GPU plumbing: the head is embarrassingly parallel over RoIs (eval-mode BN, static per-object
graphs), so one process per GPU takes a contiguous RoI shard with the weights and graphs replicated,
and the only exchange is one all-gather of the fixed-size correspondence records.
NCCL over NVLink on the GPU box; the same code runs on gloo/CPU tensors for the host-logic tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Best effort: pin the calling process to the CPUs NVML lists as local to GPU ``device_index``, so that the pinned
    host buffers it allocates afterwards (first touch) live on that GPU's NUMA node.  With one process per GPU on a
    two-socket box the host->device uploads of the end-to-end path otherwise cross the socket interconnect.  Returns
    False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import os
        import pynvml
        pynvml.nvmlInit()
        # honour CUDA_VISIBLE_DEVICES: map the torch index to the physical one when it is a plain index list
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = device_index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return True
    except Exception:
        return False


def shard_range(total: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of ``total`` RoIs for ``rank``; sizes differ by at most one."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(total: int, world: int):
    return [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]


def gather_correspondences(records: torch.Tensor, total: int | None = None, group=None) -> torch.Tensor:
    """All-gather per-rank (b_r, N, 3) int32 record shards into the full (total, N, 3) tensor on every rank.

    Equal shards use one ``all_gather_into_tensor`` (a single NCCL collective, graph-capturable); ragged
    shards are padded to the largest shard first.  Payload is 12 bytes per keypoint (48 KB per RoI at
    N=4096), so the collective is latency- not bandwidth-bound on NVLink 5.
    """
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return records
    world = dist.get_world_size(group)
    b = records.shape[0]
    total = b * world if total is None else total
    sizes = shard_sizes(total, world)
    bmax = max(sizes)
    if b != sizes[dist.get_rank(group)]:
        raise RuntimeError(f"rank holds {b} RoIs but its shard of {total} is {sizes[dist.get_rank(group)]}")
    send = records.contiguous()
    if b < bmax:
        pad = torch.zeros((bmax - b,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
        send = torch.cat([send, pad], dim=0)
    out = torch.empty((world * bmax,) + tuple(records.shape[1:]), dtype=records.dtype, device=records.device)
    try:
        dist.all_gather_into_tensor(out, send, group=group)
    except (RuntimeError, NotImplementedError):  # backends without the fused variant
        parts = [torch.empty_like(send) for _ in range(world)]
        dist.all_gather(parts, send, group=group)
        out = torch.cat(parts, dim=0)
    if all(s == bmax for s in sizes):
        return out
    out = out.view((world, bmax) + tuple(records.shape[1:]))
    return torch.cat([out[r, :sizes[r]] for r in range(world)], dim=0)


class OverlappedGather:
    """The all-gather of a step's records on a side stream, overlapped with the next step's compute.

    ``submit(records)`` orders the collective after everything enqueued so far on the current (compute) stream, runs it
    on the gather stream into one of ``depth`` rotating output buffers and returns ``(gathered, done_event)`` without
    blocking the compute stream; a consumer on another stream waits on ``done_event``.  With one rank it returns the
    records themselves.  The caller must ``torch.cuda.synchronize()`` (or wait on the events) before the results of the
    last steps count as delivered -- bench.py's timed regions end with exactly that."""

    def __init__(self, device, depth: int = 2, group=None):
        self.device, self.depth, self.group = torch.device(device), depth, group
        self.stream = torch.cuda.Stream(device=self.device)
        self._events = [None] * depth
        self._i = 0

    def submit(self, records: torch.Tensor, total: int | None = None):
        done = torch.cuda.Event()
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            done.record(torch.cuda.current_stream(self.device))
            return records, done
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            records.record_stream(self.stream)
            out = gather_correspondences(records, total, self.group)
            done.record(self.stream)
        self._events[self._i % self.depth] = done
        self._i += 1
        return out, done
