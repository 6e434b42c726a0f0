"""Multi-GPU sharding of a batch of independent QPs: contiguous slices, no collective on the solve path
(SURVEY.md section 8e).  One process per GPU; rank r solves problems [lo, hi)."""
from __future__ import annotations


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice of ceil(n / world) problems per rank (the last ranks may get fewer, possibly none)."""
    if world < 1 or not (0 <= rank < world) or n < 0:
        raise ValueError("bad shard arguments")
    per = -(-n // world)
    lo = min(n, rank * per)
    hi = min(n, lo + per)
    return lo, hi


def gather_stats(local: dict, group=None) -> dict:
    """Off-the-clock reduction of per-rank solver statistics (sums and maxima) with torch.distributed, when
    initialised; identity otherwise.  Keys ending in '_max' are max-reduced, everything else is summed."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        return dict(local)
    if not (dist.is_available() and dist.is_initialized()):
        return dict(local)
    out = {}
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    for k in sorted(local):
        t = torch.tensor([float(local[k])], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX if k.endswith("_max") else dist.ReduceOp.SUM, group=group)
        out[k] = float(t.item())
    return out
