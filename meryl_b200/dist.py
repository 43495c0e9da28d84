"""Multi-GPU plumbing around the C ABI (one process per GPU, torch.distributed for the control plane).

The data plane -- the all-to-all of packed k-mers -- is NCCL inside libmeryl_b200.so
(csrc/exchange.cu).  What the caller must provide is (1) the broadcast of the 128-byte NCCL unique
id from rank 0 and (2) a split of the input over ranks; both are here so that bench.py and the
tests share them.  Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def broadcast_unique_id(make_id, device: torch.device | str = "cpu", group=None) -> bytes:
    """Rank 0 calls make_id() -> 128 bytes; everybody gets them."""
    rank = dist.get_rank(group)
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        buf.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0, group=group)
    return bytes(buf.cpu().numpy().tobytes())


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced split of n_items: returns (first, count) for this rank."""
    base, extra = divmod(n_items, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def owned_files(rank: int, world: int) -> list[int]:
    """Data files (0..63) a rank owns: min(f, 63-f) mod world == rank (csrc/exchange.cu)."""
    return [f for f in range(64) if min(f, 63 - f) % world == rank]


def all_ranks_sum(value: int, device: torch.device | str = "cpu", group=None) -> int:
    t = torch.tensor([value], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def all_ranks_max(value: float, device: torch.device | str = "cpu", group=None) -> float:
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
