"""Multi-GPU plumbing for the hot path: independent items shard by contiguous
ranges, one process per GPU; the only collective is ONE broadcast of the 32-byte rho (shared-key
batches).  Works on any torch.distributed backend (nccl on the B200 box, gloo in the CPU tests)."""
from typing import Tuple


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous item range [lo, hi) of `rank`; ranges differ by at most one item and cover
    [0, total) exactly (GPU g gets items [g*B/G, (g+1)*B/G) when G divides B)."""
    if world < 1 or not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard arguments")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_rho(rho, src: int = 0):
    """Broadcast the 32-byte public seed rho from `src` to every rank (in place); returns rho.
    No-op when torch.distributed is not initialised (single GPU)."""
    import torch.distributed as dist
    if rho.numel() != 32:
        raise ValueError("rho must be 32 bytes")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(rho, src=src)
    return rho


def max_over_ranks(value: float, device=None) -> float:
    """Max of a per-rank timing over all ranks (how every multi-GPU number is reported)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
