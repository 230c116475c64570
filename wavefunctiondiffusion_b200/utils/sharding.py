"""Map sharding across ranks (one process per GPU). Maps of a batch are independent (the loss is per map,
reference wfd/wfdf/losses.py:69-75; GroupNorm and attention are per sample), so the data path has no collective:
ranks only meet to barrier, to take the max-over-ranks time and to gather the small results (losses, int tiles)."""
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) of `n_items` for `rank`; sizes differ by at most one, lower ranks get the extras."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def balance_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of items (e.g. maps of different tilesets, cost = F_step of the
    tileset) to `world` ranks; returns the item indices of each rank, deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (loads[k], k))
        out[r].append(i)
        loads[r] += costs[i]
    for lst in out:
        lst.sort()
    return out


def max_over_ranks(value: float, device) -> float:
    """Device-timed durations are reported as the max over ranks."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def gather_results(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All-gathers per-rank result rows (losses (b,), tiles (b, H, W)) of possibly different counts, in rank order."""
    if not (dist.is_initialized() and dist.get_world_size() > 1):
        return local
    world = dist.get_world_size()
    cap = max(counts)
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])
