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


def allreduce_gradients(model, average=True):
    """The one real exchange step of the data-parallel path: every rank holds a replica, scenes are
    sharded, and after backward the 1.06 M gradient elements (4.26 MB fp32) are summed in ONE flat bucket -- a single
    latency-bound all-reduce per step instead of one per parameter (the reference uses nn.DataParallel's reduce_add).
    BatchNorm statistics stay per replica, as with DataParallel.  Returns the number of elements reduced."""
    # EVERY trainable parameter takes part, in module order, with zeros standing in for a missing gradient: the bucket
    # layout must be identical on all ranks even when a rank's shard left a head unused (no labelled object -> no grad)
    params = [p for p in model.parameters() if p.requires_grad]
    if not params:
        return 0
    dtypes = {p.dtype for p in params}
    if len(dtypes) != 1:
        raise TypeError("allreduce_gradients: one flat bucket needs a single parameter dtype, got %s" % sorted(map(str, dtypes)))
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in params])
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        if average:
            flat /= dist.get_world_size()
    off = 0
    for p in params:
        n = p.numel()
        if p.grad is None:
            p.grad = flat[off:off + n].view_as(p).clone()
        else:
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n
    return int(flat.numel())


class GradientBuckets:
    """The gradient exchange overlapped with the tail of backward.

    Parameters are cut into `n_buckets` contiguous groups in REVERSE module order (backward reaches the heads first, the
    backbone last).  A post-accumulate-grad hook counts the gradients of each group; when a group is complete its flat
    bucket is all-reduced asynchronously (NCCL runs it on its own stream) while autograd keeps working on the earlier
    layers.  finish() waits for the outstanding reductions, averages and scatters the results back.  A parameter that
    received no gradient contributes zeros, so every rank reduces identical layouts (see allreduce_gradients)."""

    def __init__(self, model, n_buckets=3, average=True):
        self.average = average
        self.params = [p for p in model.parameters() if p.requires_grad]
        dtypes = {p.dtype for p in self.params}
        if len(dtypes) > 1:
            raise TypeError("GradientBuckets: a single parameter dtype is required, got %s" % sorted(map(str, dtypes)))
        rev = list(reversed(self.params))
        total = sum(p.numel() for p in rev)
        self.buckets, cur, acc = [], [], 0
        for p in rev:
            cur.append(p)
            acc += p.numel()
            if acc >= total * (len(self.buckets) + 1) / max(n_buckets, 1) and len(self.buckets) < n_buckets - 1:
                self.buckets.append(cur)
                cur = []
        if cur:
            self.buckets.append(cur)
        self.owner = {id(p): bi for bi, b in enumerate(self.buckets) for p in b}
        self.pending = [0] * len(self.buckets)
        self.work = [None] * len(self.buckets)
        self.flat = [None] * len(self.buckets)
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in self.params]
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        self.begin()

    def begin(self):
        """Call before every backward."""
        self.pending = [len(b) for b in self.buckets]
        self.work = [None] * len(self.buckets)
        self.flat = [None] * len(self.buckets)

    def _launch(self, bi):
        bucket = self.buckets[bi]
        self.flat[bi] = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        if self.active:
            self.work[bi] = dist.all_reduce(self.flat[bi], op=dist.ReduceOp.SUM, async_op=True)

    def _hook(self, p):
        bi = self.owner[id(p)]
        self.pending[bi] -= 1
        if self.pending[bi] == 0:
            self._launch(bi)

    def finish(self):
        """After backward: reduce the groups whose hooks never completed (unused parameters), wait, write back.
        Returns the number of elements reduced."""
        n = 0
        world = dist.get_world_size() if self.active else 1
        for bi, bucket in enumerate(self.buckets):
            if self.flat[bi] is None:
                self._launch(bi)
            if self.work[bi] is not None:
                self.work[bi].wait()
            flat = self.flat[bi]
            if self.average and world > 1:
                flat /= world
            off = 0
            for p in bucket:
                k = p.numel()
                if p.grad is None:
                    p.grad = flat[off:off + k].view_as(p).clone()
                else:
                    p.grad.copy_(flat[off:off + k].view_as(p.grad))
                off += k
            n += off
        return n

    def remove(self):
        for h in self.handles:
            h.remove()


def broadcast_parameters(model, src=0):
    """Identical replicas at start-up (parameters and buffers), one flat broadcast."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    tensors = [t for t in list(model.parameters()) + list(model.buffers()) if t.is_floating_point()]
    flat = torch.cat([t.detach().reshape(-1) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    with torch.no_grad():
        for t in tensors:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
