"""Batch sharding across ranks (one process per GPU, torch.distributed plumbing only).

The GPAD batch partitions into independent QPs, so there is NO collective inside the iteration
loop: every rank solves its own contiguous shard with its own replicated copy of the operators.
The only exchange of the path is the final gather of the first control move u0 = z[:, :n_u]
(what the MPC applies, gpad.m:91) and the per-instance status, plus the max-over-ranks of the
device time for reporting."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """contiguous [lo, hi) of `total` instances owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_seed(base_seed, rank):
    """per-rank seed of the synthetic workload (weak scaling: every rank generates its own shard)"""
    return base_seed + rank


def gather_first_moves(z, n_u, dst=0, equal_shards=False):
    """gather u0 = z[:, :n_u] of every rank on `dst`; returns [sum of shard sizes, n_u] there, None
    elsewhere.  equal_shards=True skips the (host-synchronising) exchange of shard sizes."""
    u0 = z[:, :n_u].contiguous()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return u0
    world, rank = dist.get_world_size(), dist.get_rank()
    if equal_shards:
        sizes = [u0.shape[0]] * world
    else:
        sizes = [torch.zeros(1, dtype=torch.int64, device=u0.device) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([u0.shape[0]], dtype=torch.int64, device=u0.device))
        sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1:
        out = [torch.empty_like(u0) for _ in range(world)] if rank == dst else None
        dist.gather(u0, out, dst=dst)
        return torch.cat(out, 0) if rank == dst else None
    # ragged shards: pad to the largest, gather, trim
    cap = max(sizes)
    padded = torch.zeros((cap, n_u), dtype=u0.dtype, device=u0.device)
    padded[:u0.shape[0]] = u0
    out = [torch.empty_like(padded) for _ in range(world)] if rank == dst else None
    dist.gather(padded, out, dst=dst)
    return torch.cat([o[:s] for o, s in zip(out, sizes)], 0) if rank == dst else None


def max_over_ranks(value, device="cpu"):
    """max of a python float over all ranks (device time of the slowest rank)"""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
