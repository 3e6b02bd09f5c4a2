"""Multi-GPU sharding of a problem batch: problems are independent, so each rank solves a contiguous
block with the identical kernel sequence and there is NO collective on the solve path; only per-problem result
scalars (cost, status, iteration counts) are gathered afterwards (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(B_total, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; blocks differ by at most one problem."""
    base, rem = divmod(int(B_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_scalars(local, B_total, group=None):
    """All-gather a per-problem scalar tensor [B_local] (any dtype) into the global [B_total] on every rank.
    Ragged shards (B_total not divisible by the world size) are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local.clone()
    world = dist.get_world_size(group)
    sizes = [shard_range(B_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(mx, dtype=local.dtype, device=local.device)
    buf[:local.shape[0]] = local
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)])
