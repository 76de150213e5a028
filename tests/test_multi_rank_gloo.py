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
