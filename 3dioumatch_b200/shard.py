"""Scene sharding and max-over-ranks timing for the one-process-per-GPU harness.

The hot path has no cross-scene data flow (every kernel is per scene), so multi-GPU execution is pure scene sharding:
rank r of W owns scenes [r*ceil(S/W), ...) of a global batch, runs them independently, and only the timing / the
result gathering use a collective.  Works on any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def scene_shard(rank, world, total):
    """Contiguous, balanced shard of `total` scenes: the first (total % world) ranks own one extra scene."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value, device="cpu"):
    """Elapsed time of a step = the slowest rank's time."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_scenes(local, total, dim=0):
    """All-gather per-scene results (tensor with `dim` = local scenes) back into global scene order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [scene_shard(r, world, total) for r in range(world)]
    pad = max(e - s for s, e in sizes)
    shape = list(local.shape)
    shape[dim] = pad
    buf = torch.zeros(shape, dtype=local.dtype, device=local.device)
    buf.narrow(dim, 0, local.shape[dim]).copy_(local)
    outs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(outs, buf)
    return torch.cat([o.narrow(dim, 0, e - s) for o, (s, e) in zip(outs, sizes)], dim=dim)


def throughput(total_units, elapsed_ms):
    return total_units / (elapsed_ms / 1e3)
