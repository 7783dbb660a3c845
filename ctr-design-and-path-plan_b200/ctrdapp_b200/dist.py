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
    """Host-side convenience over reduce_best: every rank passes its own (cost, index), all get the global winner
    (lowest cost, then lowest index)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(best_cost), int(best_index)
    world = dist.get_world_size(group)
    # the index travels as an integer (a float64 payload is exact only below 2^53)
    key = cost_to_key(float(best_cost)) if int(best_index) >= 0 else INT64_MAX
    t = torch.tensor([key, int(best_index) if int(best_index) >= 0 else INT64_MAX], dtype=torch.int64, device=device)
    return decode_best(reduce_best(t, group=group))


INT64_MAX = (1 << 63) - 1


def cost_to_key(cost):
    """Order-preserving signed 64-bit image of a cost (ctrb_cost_to_key, in NumPy): key(a) < key(b) iff a < b; NaN and
    +inf (colliding / no candidate) map above every finite cost."""
    c = np.asarray(cost, np.float64)
    b = c.view(np.uint64) if c.ndim else np.array([c]).view(np.uint64)
    neg = (b >> np.uint64(63)).astype(bool)
    u = np.where(neg, ~b, b | np.uint64(1 << 63))
    k = (u ^ np.uint64(1 << 63)).view(np.int64)
    k = np.where(np.isfinite(c.reshape(k.shape)), k, INT64_MAX)
    return k if c.ndim else int(k[0])


def key_to_cost(key):
    key = int(key)
    if key == INT64_MAX:
        return float("inf")
    b = np.uint64(key & 0xFFFFFFFFFFFFFFFF) ^ np.uint64(1 << 63)
    b = (b & np.uint64(0x7FFFFFFFFFFFFFFF)) if (b >> np.uint64(63)) else ~b
    return float(np.array([b], np.uint64).view(np.float64)[0])


def reduce_best(key_index, group=None):
    """The best node over all ranks without leaving the device: key_index is this rank's int64 tensor
    [key, global index] (ctrb_argmin_cost_device, or [cost_to_key(c), i] on the host).  Two 8-byte all-reduce(min)
    steps: the cost key, then the index among the ranks that hold that key (lowest cost, then lowest index -- exact,
    where one packed word would have to truncate either the cost or the index).  In place; returns the tensor."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return key_index
    mine = key_index[0:1].clone()
    dist.all_reduce(key_index[0:1], op=dist.ReduceOp.MIN, group=group)
    key_index[1:2] = torch.where(mine == key_index[0:1], key_index[1:2], torch.full_like(mine, INT64_MAX))
    dist.all_reduce(key_index[1:2], op=dist.ReduceOp.MIN, group=group)
    return key_index


def decode_best(key_index):
    """(cost, index) of a reduced [key, index] tensor (one device-to-host read); (inf, -1) if no rank had a candidate."""
    k, i = (int(v) for v in key_index.cpu().tolist())
    if k == INT64_MAX or i == INT64_MAX:
        return float("inf"), -1
    return key_to_cost(k), i


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
