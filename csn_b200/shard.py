"""Multi-GPU partitioning of the hot path: one process per GPU, shapes sharded by
QUERY shape, no collective inside a step's data path.

* retrieval: rank r scores queries shard_range(S, r, W) against the replicated candidate store and
  the (S/W, K+1) index rows are gathered once at the end (gather_rows);
* training: data-parallel over query shapes, one flat all-reduce of the parameter gradients per step
  (allreduce_mean_).
Backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of [0, n) owned by `rank`; block sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local: torch.Tensor, n_total: int, rank: int, world: int) -> torch.Tensor:
    """All ranks receive the (n_total, ...) tensor whose row blocks are the ranks' `local` tensors."""
    if world == 1:
        return local
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    pad = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((pad,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)


def allreduce_mean_(tensors, world: int) -> None:
    """In-place mean over ranks of a list of tensors through ONE flat all-reduce (0.64 M floats at
    h = 1: a single bucket, sized for launch latency rather than link count)."""
    tensors = [t for t in tensors if t is not None]
    if world == 1 or not tensors:
        return
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat)
    flat.div_(world)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
