"""Multi-GPU plumbing: test points shard trivially, one process per GPU.

Each test point's outputs depend only on that point and the small read-only model (reference
gp_emulator/GaussianProcess.py:228-249), so rank r of G owns a contiguous row range and there is no
data-path collective.  The only communication is a one-time broadcast of the trained model from rank 0
(<= 0.5 MB at M = 250) and the max-over-ranks reduction of timings.  Works with the ``nccl`` backend on
GPUs and with ``gloo`` on CPU tensors (used by the world_size-2 CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(N, rank, world):
    """Contiguous [start, end) of N rows owned by ``rank``; sizes differ by at most one row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    base, rem = divmod(int(N), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def bind_host_to_gpu(device_index):
    """Pin this process to the CPUs that are NUMA-local to GPU ``device_index`` (NVML's CPU affinity mask).

    With one process per GPU every rank streams its own test points host -> device and results back; page-locked
    buffers are placed by first touch, so a rank that runs on the far socket pushes all of its PCIe traffic across
    the inter-socket link.  Call this before allocating host buffers.  Returns the CPU list, or None when NVML or
    the affinity call is unavailable (nothing is changed then).
    """
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
            ncpu = os.cpu_count() or 1
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        finally:
            pynvml.nvmlShutdown()
        cpus = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = sorted(c for c in cpus if c in allowed)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def broadcast_model(model, src=0, device=None):
    """Broadcast a dict of float64 numpy arrays (the trained model) from ``src`` to every rank.

    Non-source ranks pass ``None``.  Shapes travel first (object broadcast), then one flat payload, so the
    transfer is a single collective regardless of how many arrays the model has.
    """
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    meta = [None]
    if rank == src:
        meta[0] = [(k, tuple(np.shape(v))) for k, v in model.items()]
    dist.broadcast_object_list(meta, src=src)
    total = int(sum(int(np.prod(s)) for _, s in meta[0]))
    if rank == src:
        flat = torch.from_numpy(np.concatenate([np.asarray(model[k], dtype=np.float64).ravel() for k, _ in meta[0]]))
    else:
        flat = torch.empty(total, dtype=torch.float64)
    if device is not None:
        flat = flat.to(device)
    dist.broadcast(flat, src=src)
    flat = flat.cpu().numpy()
    out, off = {}, 0
    for k, s in meta[0]:
        n = int(np.prod(s))
        out[k] = flat[off:off + n].reshape(s).copy()
        off += n
    return out


def max_over_ranks(value, device=None):
    """All-reduce MAX of a python float (per-rank device time -> job time)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
