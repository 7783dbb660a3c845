"""Multi-GPU host logic: one process per GPU (torchrun), configurations sharded by contiguous blocks,
no data-path collective; the only exchange is each rank's best node (16 bytes) and, when the planner
wants them, the per-configuration costs.  Backend-agnostic (`nccl` on the GPUs, `gloo` in
the CPU tests)."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` configurations owned by `rank` (sizes differ by at most 1)."""
    base, rem = divmod(int(total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def local_best(cost, index_base=0):
    """(cost, global index) of the lowest finite cost of a host array; (inf, -1) if none (NaN = discarded)."""
    cost = np.asarray(cost, np.float64)
    finite = np.isfinite(cost)
    if not finite.any():
        return float("inf"), -1
    k = int(np.argmin(np.where(finite, cost, np.inf)))
    return float(cost[k]), index_base + k


def gather_best(best_cost, best_index, device="cpu", group=None):
    """all_gather of every rank's (cost, index); returns the global winner (lowest cost, then lowest index)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(best_cost), int(best_index)
    world = dist.get_world_size(group)
    mine = torch.tensor([float(best_cost), float(best_index)], dtype=torch.float64, device=device)
    out = torch.empty(world * 2, dtype=torch.float64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    g = out.view(world, 2).cpu().numpy()
    valid = g[:, 1] >= 0
    if not valid.any():
        return float("inf"), -1
    order = np.lexsort((g[:, 1], np.where(valid, g[:, 0], np.inf)))
    return float(g[order[0], 0]), int(g[order[0], 1])


def gather_costs(cost_local, device="cpu", group=None):
    """all_gather of equally sized per-configuration cost shards -> the whole batch's costs on every rank."""
    import torch
    import torch.distributed as dist
    t = cost_local if isinstance(cost_local, torch.Tensor) else \
        torch.as_tensor(np.ascontiguousarray(cost_local, np.float64), device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    out = torch.empty(dist.get_world_size(group) * t.numel(), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out
