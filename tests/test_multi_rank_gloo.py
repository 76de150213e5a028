"""CPU: the N>1 path of the harness (scene sharding, max-over-ranks timing, result gathering) on a world_size-2 gloo group.
The per-scene work is done with the CPU oracle here (no GPU in this suite); the GPU run uses the same shard logic with
the sm_100a kernels (bench.py under torchrun)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("3dioumatch_b200.shard")
    import cases
    import oracle as orc
    s, e = shard.scene_shard(rank, world, total)
    xyz = cases.cloud(0, total, 700)                      # the same global batch on every rank
    local = torch.from_numpy(orc.furthest_point_sampling(xyz[s:e], 32).astype(np.int64))
    full = shard.gather_scenes(local, total)
    slowest = shard.max_over_ranks(10.0 + rank)           # rank 1 is "slower"
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), full.numpy())
        np.save(os.path.join(out_dir, "slowest.npy"), np.asarray([slowest]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_partition_properties():
    shard = importlib.import_module("3dioumatch_b200.shard")
    for total in (1, 7, 8, 9, 64):
        for world in (1, 2, 3, 8):
            spans = [shard.scene_shard(r, world, total) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))             # contiguous, no overlap
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1                                     # balanced


@pytest.mark.timeout(120)
def test_two_rank_gloo_sharded_run_matches_single_process(tmp_path, orc):
    import cases
    total = 5                                                                       # uneven shards: 3 + 2
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, total, str(tmp_path)), nprocs=2, join=True)
    gathered = np.load(tmp_path / "gathered.npy")
    ref = orc.furthest_point_sampling(cases.cloud(0, total, 700), 32)
    assert np.array_equal(gathered, ref)                                            # every scene once, in order
    assert float(np.load(tmp_path / "slowest.npy")[0]) == 11.0                      # max over ranks


def _grad_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("3dioumatch_b200.shard")
    torch.manual_seed(100 + rank)                      # replicas start different on purpose
    net = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 1), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Conv1d(8, 2, 1))
    shard.broadcast_parameters(net, src=0)
    torch.manual_seed(0)
    x = torch.randn(6, 4, 10)                          # global batch of 6 "scenes"
    s, e = shard.scene_shard(rank, world, 6)
    loss = net(x[s:e]).square().sum() / 6.0            # per-rank share of the global mean loss
    loss.backward()
    n = shard.allreduce_gradients(net, average=False)
    if rank == 0:
        torch.save({"grads": [p.grad.clone() for p in net.parameters()], "n": n,
                    "state": {k: v.clone() for k, v in net.state_dict().items()}}, os.path.join(out_dir, "g.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_bucket_gradient_allreduce(tmp_path):
    """Sharded backward + one flat all-reduce == single-process backward over the whole batch (conv grads; BatchNorm
    statistics are per replica by design, so the check uses eval-mode-free layers' gradients with BN in train mode on
    equal shard sizes, where the per-shard batch statistics differ -> compare against the same sharded computation)."""
    port = 29000 + ((os.getpid() + 7) % 2000)
    mp.spawn(_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(tmp_path / "g.pt")
    # single-process emulation of the two replicas (same broadcast weights, same shards)
    net = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 1), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Conv1d(8, 2, 1))
    torch.manual_seed(0)
    x = torch.randn(6, 4, 10)
    total = None
    for s, e in ((0, 3), (3, 6)):
        rep = torch.nn.Sequential(torch.nn.Conv1d(4, 8, 1), torch.nn.BatchNorm1d(8), torch.nn.ReLU(), torch.nn.Conv1d(8, 2, 1))
        rep.load_state_dict({k: v for k, v in got["state"].items()})
        rep[1].running_mean.zero_(); rep[1].running_var.fill_(1.0); rep[1].num_batches_tracked.zero_()
        (rep(x[s:e]).square().sum() / 6.0).backward()
        g = [p.grad for p in rep.parameters()]
        total = g if total is None else [a + b for a, b in zip(total, g)]
    assert got["n"] == sum(p.numel() for p in net.parameters())
    for a, b in zip(got["grads"], total):
        assert torch.allclose(a, b, atol=1e-6), (a - b).abs().max()


def _bucket_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shard = importlib.import_module("3dioumatch_b200.shard")
    torch.manual_seed(7)
    net = torch.nn.ModuleDict({"a": torch.nn.Linear(5, 7), "b": torch.nn.Linear(7, 3), "unused": torch.nn.Linear(3, 2),
                               "rank1_only": torch.nn.Linear(3, 1)})
    buckets = shard.GradientBuckets(net, n_buckets=3, average=True)
    torch.manual_seed(rank)
    x = torch.randn(4, 5)
    for _ in range(2):                                     # two steps: begin() must re-arm the hooks
        for p in net.parameters():
            p.grad = None
        buckets.begin()
        y = net["b"](torch.relu(net["a"](x)))
        loss = y.square().sum()
        if rank == 1:                                      # a head only this rank's shard exercises
            loss = loss + net["rank1_only"](y).sum()
        loss.backward()
        n = buckets.finish()
    torch.save({"grads": {k: p.grad.clone() for k, p in net.named_parameters()}, "n": n, "x": x,
                "state": {k: v.clone() for k, v in net.state_dict().items()}}, os.path.join(out_dir, "b%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gradient_buckets_overlap_hooks_and_unused_parameters(tmp_path):
    """GradientBuckets (hook-launched asynchronous bucket all-reduces): every rank ends with the MEAN gradient, parameters
    without a gradient on some (or all) ranks contribute zeros instead of desynchronising the bucket layout."""
    port = 29000 + ((os.getpid() + 13) % 2000)
    mp.spawn(_bucket_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "b0.pt"), torch.load(tmp_path / "b1.pt")
    assert r0["n"] == r1["n"] == sum(v.numel() for v in r0["state"].values())
    for k in r0["grads"]:
        assert torch.equal(r0["grads"][k], r1["grads"][k]), k          # identical on both ranks after the exchange
    # single-process reference
    ref = {}
    for rank, rec in enumerate((r0, r1)):
        net = torch.nn.ModuleDict({"a": torch.nn.Linear(5, 7), "b": torch.nn.Linear(7, 3), "unused": torch.nn.Linear(3, 2),
                                   "rank1_only": torch.nn.Linear(3, 1)})
        net.load_state_dict(rec["state"])
        y = net["b"](torch.relu(net["a"](rec["x"])))
        loss = y.square().sum()
        if rank == 1:
            loss = loss + net["rank1_only"](y).sum()
        loss.backward()
        for k, p in net.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            ref[k] = ref.get(k, 0) + g / 2
    for k in ref:
        assert torch.allclose(r0["grads"][k], ref[k], atol=1e-6), k
    assert float(r0["grads"]["unused.weight"].abs().sum()) == 0.0
