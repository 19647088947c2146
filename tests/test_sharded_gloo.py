"""N>1 host logic on CPU: world_size-2 gloo processes exercise the shard exchange
(all-gather layout, global id packing, per-shard limit); the merge semantics are checked
against the oracle (cluster/actions.go:291-299,357-376). No GPU needed."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oraclelib as O
from semadb_b200 import sharded


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.Generator(np.random.PCG64(100 + rank))
    d = np.sort(rng.integers(0, 50, size=(B, k)).astype(np.float32), axis=1)
    ids = rng.integers(2, 1 << 30, size=(B, k)).astype(np.int64)
    cnt = rng.integers(0, k + 1, size=(B,)).astype(np.int32)
    g_ids, g_d, g_c = sharded.exchange_topk(sharded.pack_global_ids(torch.from_numpy(ids), rank),
                                            torch.from_numpy(d), torch.from_numpy(cnt))
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), ids=ids, d=d, cnt=cnt, g_ids=g_ids.numpy(), g_d=g_d.numpy(),
             g_c=g_c.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_exchange_topk_world2(tmp_path):
    world, B, k = 2, 64, 10
    port = _free_port()
    mp.spawn(_worker, args=(world, port, B, k, str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f"r{i}.npz") for i in range(world)]
    for me in range(world):
        assert r[me]["g_ids"].shape == (world, B, k) and r[me]["g_c"].shape == (world, B)
        for s in range(world):
            shard, local = sharded.unpack_global_ids(r[me]["g_ids"][s])
            assert (shard == s).all()
            assert (local == r[s]["ids"]).all()
            assert (r[me]["g_d"][s] == r[s]["d"]).all() and (r[me]["g_c"][s] == r[s]["cnt"]).all()
    # every rank holds the same gathered lists => the merge (K6, checked on the GPU in
    # test_distance_family.py) is replicated; oracle merge of the gathered data:
    oi, od, oc = O.merge_topk(r[0]["g_ids"].astype(np.uint64), r[0]["g_d"], r[0]["g_c"].astype(np.uint32), k)
    oi1, od1, oc1 = O.merge_topk(r[1]["g_ids"].astype(np.uint64), r[1]["g_d"], r[1]["g_c"].astype(np.uint32), k)
    assert (oi == oi1).all() and (od == od1).all() and (oc == oc1).all()
    assert (np.diff(np.where(np.isinf(od), np.float32(1e30), od), axis=1) >= 0).all()
    assert (oc == np.minimum(r[0]["g_c"].sum(axis=0), k)).all()


def test_partition_is_balanced_and_deterministic():
    a = sharded.partition_points(100_003, 8, seed=5)
    b = sharded.partition_points(100_003, 8, seed=5)
    assert (a == b).all()
    counts = np.bincount(a, minlength=8)
    assert counts.max() - counts.min() <= 1
    assert (sharded.partition_points(100_003, 8, seed=6) != a).any()


def test_shard_limit_matches_reference_formula():
    # cluster/actions.go:291-299
    for limit in (1, 10, 50, 75, 100):
        for s in (1, 2, 4, 8, 16):
            assert sharded.shard_limit(limit, s) == O.shard_limit(limit, s)
    assert sharded.shard_limit(10, 8) == 10 and sharded.shard_limit(100, 5) == 38


class _FakeIndex:
    """Stands in for the device index on CPU: 'searches' by writing a function of the query."""

    def search_batch_device(self, q, k, L, ids, d, c, stream=0):
        n = q.shape[0]
        ids[:n] = (q[:, :1] * 1000).long() + torch.arange(k)[None, :]
        d[:n] = q[:, :1] + torch.arange(k, dtype=torch.float32)[None, :]
        c[:n] = k


def _replica_worker(rank, world, port, B, k, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q = torch.arange(B, dtype=torch.float32)[:, None].repeat(1, 4)
    s = sharded.ReplicatedSearcher(_FakeIndex(), rank, world)
    ids, d, c = s.search_batch_device(q, k, 75)
    np.savez(os.path.join(out_dir, f"rep{rank}.npz"), ids=ids.numpy(), d=d.numpy(), c=c.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [64, 7, 1])
def test_replicated_mode_world2_concatenates_slices_in_batch_order(tmp_path, B):
    """Replicated mode: each rank serves its contiguous slice, every rank ends with the full batch
    in order — including ragged (7) and smaller-than-world (1) batches."""
    world, k = 2, 10
    mp.spawn(_replica_worker, args=(world, _free_port(), B, k, str(tmp_path)), nprocs=world, join=True)
    want_ids = (np.arange(B)[:, None] * 1000 + np.arange(k)[None, :]).astype(np.int64)
    for r in range(world):
        z = np.load(tmp_path / f"rep{r}.npz")
        assert z["ids"].shape == (B, k) and (z["ids"] == want_ids).all()
        assert (z["d"][:, 0] == np.arange(B)).all() and (z["c"] == k).all()
    lo, hi = sharded.replica_slice(7, 0, 2)
    assert (lo, hi) == (0, 4) and sharded.replica_slice(7, 1, 2) == (4, 7) and sharded.replica_slice(1, 1, 2) == (1, 1)
