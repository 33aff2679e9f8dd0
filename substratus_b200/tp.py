"""Tensor-parallel bootstrap for the one-process-per-GPU launch (torchrun, or one serve process per GPU).

The engine's allreduce runs over NVLink peer memory (CUDA IPC); the only thing the host has to do is gather the
256-byte handles every rank exports (``ssb_tp_export``) and hand the rank-major concatenation back
(``ssb_tp_connect``).  ``torch.distributed`` is used purely as that plumbing (gloo on CPU tensors or NCCL on CUDA
tensors, whichever backend the process group has).  The reference has no multi-GPU path at all
(internal/controller/server_controller.go:115 hard-codes replicas=1, one container); see SURVEY.md §8e.
"""
from __future__ import annotations

import numpy as np


def gather_blobs(blob: bytes, group=None) -> bytes:
    """All-gather equal-length byte strings across the process group, rank-major."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return b"".join(bytes(t.cpu().numpy().tobytes()) for t in out)


def connect(engine, group=None) -> None:
    """Export this rank's TP handle, all-gather, connect.  Collective: every rank must call it."""
    blob = engine.tp_export()
    engine.tp_connect(gather_blobs(blob, group))


def shard_ranges(total: int, tp_size: int):
    """[start, end) of each rank's contiguous shard (column-parallel rows / row-parallel K-slices / KV heads)."""
    if total % tp_size:
        raise ValueError(f"{total} not divisible by tp_size {tp_size}")
    n = total // tp_size
    return [(r * n, (r + 1) * n) for r in range(tp_size)]


def max_over_ranks(value: float, group=None) -> float:
    """MAX reduction of a host scalar (bench timing: the slowest rank defines the step)."""
    import torch
    import torch.distributed as dist

    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
