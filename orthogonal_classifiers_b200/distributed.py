"""Data-parallel sharding of the encode + classify path.

Images are independent units, so the path shards with NO data-path collective:
each rank takes a contiguous slice of the images (and labels), the classifier
(<= 72 KB) is replicated, and the only exchange is one all-reduce(sum) of two
int64 — (correct, total) — per evaluation.  One process per
GPU (torchrun); NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of n items: the first n % world ranks get one
    extra.  Every item belongs to exactly one rank; empty shards are legal."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def rank_world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_reduce_counts(counts: torch.Tensor) -> torch.Tensor:
    """Sum the int64[2] = (correct, total) tensor over ranks in place (16 bytes
    on the wire).  No-op without an initialised process group."""
    if counts.dtype != torch.int64 or counts.numel() != 2:
        raise ValueError("counts must be int64[2]")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    return counts


def accuracy_from_counts(counts: torch.Tensor) -> float:
    c = counts.detach().cpu()
    return float(c[0].item()) / max(1, int(c[1].item()))


def gather_predictions(local: torch.Tensor, n_total: int) -> torch.Tensor | None:
    """Optional: rank 0 receives every rank's (n_local, S) slab in shard order
    (not part of the timed path; predictions normally stay rank-local)."""
    rank, world = rank_world()
    if world == 1:
        return local
    sizes = [shard_bounds(n_total, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)  # dist.gather needs equal shapes: pad, then trim
    padded = torch.zeros((maxn,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    bufs = [torch.empty_like(padded) for _ in range(world)] if rank == 0 else None
    dist.gather(padded, bufs, dst=0)
    if rank != 0:
        return None
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, sizes)], 0)
