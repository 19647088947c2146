"""Sharded search on the GPU: S independent shards searched by K1, lists merged by K6,
compared with the oracle doing the same (cluster/actions.go:316-376)."""
import numpy as np
import pytest

from oracle import oraclelib as O
from semadb_b200 import _capi, sharded, synth
from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters

pytestmark = pytest.mark.gpu


def test_sharded_search_matches_oracle_single_gpu():
    import torch
    from tests.helpers import mirror_to_gpu, oracle_graph
    S, n, dim, B, k = 4, 6000, 128, 256, 10
    X = synth.sift_shaped(n, dim, 3)
    Q = synth.sift_shaped(B, dim, 4, w_seed=3)
    part = sharded.partition_points(n, S, seed=1)
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(Q).to(dev)
    g_ids = torch.zeros((S, B, k), dtype=torch.int64, device=dev)
    g_d = torch.zeros((S, B, k), dtype=torch.float32, device=dev)
    g_c = torch.zeros((S, B), dtype=torch.int32, device=dev)
    o_ids = np.zeros((S, B, k), np.uint64)
    o_d = np.zeros((S, B, k), np.float32)
    o_c = np.zeros((S, B), np.uint32)
    keep = []
    for s in range(S):
        Xs = X[part == s]
        oix, ids, start = oracle_graph(Xs, start_seed=50 + s)
        g = mirror_to_gpu(oix, Xs, ids, start)
        keep.append(g)
        g.search_batch_device(d_q, k, 75, g_ids[s], g_d[s], g_c[s], torch.cuda.current_stream().cuda_stream)
        g_ids[s] = sharded.pack_global_ids(g_ids[s], s)
        ref = oix.search(Q, k=k, threads=8)
        o_ids[s] = sharded.pack_global_ids(ref["ids"].astype(np.uint64), s)
        o_d[s] = ref["dists"]
        o_c[s] = ref["counts"]
    torch.cuda.synchronize()
    assert (g_ids.cpu().numpy().astype(np.uint64) == o_ids).all()
    m_ids = torch.zeros((B, k), dtype=torch.int64, device=dev)
    m_d = torch.zeros((B, k), dtype=torch.float32, device=dev)
    m_c = torch.zeros((B,), dtype=torch.int32, device=dev)
    _capi.check(_capi.lib().sdb_merge_topk_device(0, S, B, k, g_ids.data_ptr(), g_d.data_ptr(), g_c.data_ptr(),
                                                  m_ids.data_ptr(), m_d.data_ptr(), m_c.data_ptr(),
                                                  torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    oi, od, oc = O.merge_topk(o_ids, o_d, o_c, k)
    assert (m_ids.cpu().numpy().astype(np.uint64) == oi).all()
    assert m_d.cpu().numpy().tobytes() == od.tobytes()
    assert (m_c.cpu().numpy() == oc).all()
    # merged result = exact top-k of the union of per-shard lists; recall vs global brute force
    full = O.OracleIndex(dim)
    full.set_vectors(np.arange(2, n + 2, dtype=np.uint32), X)
    gt = full.flat_search(Q, k=k, threads=8)
    # map global ids back to original rows
    rows_of = [np.nonzero(part == s)[0] for s in range(S)]
    shard, local = sharded.unpack_global_ids(oi.astype(np.int64))
    orig = np.zeros_like(local)
    for s in range(S):
        m = shard == s
        orig[m] = rows_of[s][local[m] - 2] + 2
    rec = np.mean([len(set(orig[b].tolist()) & set(gt["ids"][b].tolist())) / k for b in range(B)])
    assert rec >= 0.99


def test_fused_gather_matches_unfused_single_gpu():
    """sdb_search_batch_gather_device: the epilogue stores tagged top-k lists straight into the
    gather buffer slot [shard]; with every shard on one GPU the buffer must equal what the
    separate search + pack produces, and barrier + merge must give the same merged lists."""
    import ctypes as C

    import torch
    from tests.helpers import mirror_to_gpu, oracle_graph
    S, n, dim, B, k = 3, 3000, 128, 200, 10
    X = synth.sift_shaped(n, dim, 3)
    Q = synth.sift_shaped(B, dim, 4, w_seed=3)
    part = sharded.partition_points(n, S, seed=2)
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream
    d_q = torch.from_numpy(Q).to(dev)
    lib = _capi.lib()
    u_ids = torch.zeros((S, B, k), dtype=torch.int64, device=dev)
    u_d = torch.zeros((S, B, k), dtype=torch.float32, device=dev)
    u_c = torch.zeros((S, B), dtype=torch.int32, device=dev)
    f_ids = torch.full((S, B, k), -1, dtype=torch.int64, device=dev)
    f_d = torch.full((S, B, k), -1.0, dtype=torch.float32, device=dev)
    f_c = torch.full((S, B), -1, dtype=torch.int32, device=dev)
    flags = torch.zeros((64,), dtype=torch.int32, device=dev)
    l_ids = torch.zeros((B, k), dtype=torch.int64, device=dev)
    l_d = torch.zeros((B, k), dtype=torch.float32, device=dev)
    l_c = torch.zeros((B,), dtype=torch.int32, device=dev)
    keep = []
    for s in range(S):
        Xs = X[part == s]
        oix, ids, start = oracle_graph(Xs, start_seed=70 + s)
        g = mirror_to_gpu(oix, Xs, ids, start)
        keep.append(g)
        g.search_batch_device(d_q, k, 75, u_ids[s], u_d[s], u_c[s], st)
        u_ids[s] = sharded.pack_global_ids(u_ids[s], s)
        pg = _capi.SdbPeerGather()
        pg.n_peers, pg.shard, pg.per_shard_limit = 1, s, 0
        pg.ids[0], pg.dists[0], pg.counts[0] = f_ids.data_ptr(), f_d.data_ptr(), f_c.data_ptr()
        _capi.check(lib.sdb_search_batch_gather_device(g._h, B, d_q.data_ptr(), k, 75, l_ids.data_ptr(), l_d.data_ptr(),
                                                       l_c.data_ptr(), C.byref(pg), st))
    pf = (C.c_void_p * 1)(flags.data_ptr())
    _capi.check(lib.sdb_peer_barrier_device(0, 1, 0, pf, 1, st))
    torch.cuda.synchronize()
    cnt = u_c.cpu().numpy()
    assert (f_c.cpu().numpy() == cnt).all()
    assert (f_ids.cpu().numpy() == u_ids.cpu().numpy()).all()
    assert f_d.cpu().numpy().tobytes() == u_d.cpu().numpy().tobytes()
    assert int(flags[0].item()) == 1 and int(flags[_capi.MAX_PEERS].item()) == 0
    # per-shard limit below k clamps the published counts only (cluster/actions.go:291-299)
    pg.per_shard_limit = 4
    _capi.check(lib.sdb_search_batch_gather_device(keep[-1]._h, B, d_q.data_ptr(), k, 75, l_ids.data_ptr(),
                                                   l_d.data_ptr(), l_c.data_ptr(), C.byref(pg), st))
    torch.cuda.synchronize()
    assert (f_c[S - 1].cpu().numpy() == np.minimum(cnt[S - 1], 4)).all()
    assert (l_c.cpu().numpy() == cnt[S - 1]).all()


def _p2p_worker(rank, world, port, out_dir):
    import os

    import torch
    import torch.distributed as dist
    from tests.helpers import mirror_to_gpu, oracle_graph
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n, dim, B, k = 4000, 128, 300, 10
    X = synth.sift_shaped(n, dim, 3 + rank)
    Q = synth.sift_shaped(B, dim, 4, w_seed=3)
    oix, ids, start = oracle_graph(X, start_seed=80 + rank)
    params = IndexVectorVamanaParameters(dim, "euclidean", 75, 64, 1.2)
    g = IndexVamana("p2p", params, device=rank, start_vector=start)
    g.set_vectors(ids.astype(np.uint64), X)
    adj, deg = oix.get_graph()
    g.set_graph_dense(adj[1:], deg[1:], first_id=1)
    d_q = torch.from_numpy(Q).to(dev)
    a = sharded.ShardedSearcher(g, rank, world, exchange="nccl")
    ref = [t.clone() for t in a.search_batch_device(d_q, k, 75)]
    b = sharded.ShardedSearcher(g, rank, world, exchange="p2p")
    same = True
    for _ in range(5):  # both buffer parities, growing epochs
        got = b.search_batch_device(d_q, k, 75)
        torch.cuda.synchronize()
        same = same and all(bool((x == y).all().item()) for x, y in zip(got, ref))
    # host-facing twin: pinned buffers read / written in place by the kernels
    h_q = torch.from_numpy(Q).pin_memory()
    h_i = torch.zeros((B, k), dtype=torch.int64).pin_memory()
    h_d = torch.zeros((B, k), dtype=torch.float32).pin_memory()
    h_c = torch.zeros((B,), dtype=torch.int32).pin_memory()
    for _ in range(2):
        b.search_batch_pinned(h_q, k, 75, h_i, h_d, h_c, dev)
        same = same and bool((h_i == ref[0].cpu()).all()) and bool((h_d == ref[1].cpu()).all()) and \
            bool((h_c == ref[2].cpu()).all())
    # pipelined steps: the exchange of step e overlaps the search of step e+1 (three gather
    # buffers, side stream). Different query batches per step, submitted back to back: every
    # step's lists must equal the unpipelined answer for its batch.
    c = sharded.ShardedSearcher(g, rank, world, exchange="p2p", pipeline=True)
    batches = [torch.roll(d_q, shifts=7 * i, dims=0).contiguous() for i in range(7)]
    want = []
    for qb in batches:
        want.append([t.clone() for t in a.search_batch_device(qb, k, 75)])
    torch.cuda.synchronize()
    got_p = []
    for i, qb in enumerate(batches):
        out = c.search_batch_device(qb, k, 75)
        if i >= 2:  # results of step i-2 are about to lose their buffer: take them (stream-ordered)
            c.wait_pipeline()
        got_p.append([t.clone() for t in out])  # clone is ordered after wait_pipeline on the current stream only from i >= 2
    c.wait_pipeline()
    torch.cuda.synchronize()
    final = [t.clone() for t in out]
    same_pipe = all(bool((x == y).all().item()) for x, y in zip(final, want[-1]))
    for i in range(2, len(batches)):
        same_pipe = same_pipe and all(bool((x == y).all().item()) for x, y in zip(got_p[i], want[i]))
    failed = b._peer.barrier_failed() or c._peer.barrier_failed()
    np.savez(os.path.join(out_dir, f"p2p{rank}.npz"), same=same and same_pipe, failed=failed, ids=ref[0].cpu().numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_fused_gather_two_gpus_matches_nccl_path(tmp_path):
    """world_size 2, one process per GPU: peer-mapped gather buffers + flag barrier give the
    same merged lists on every rank as the all-gather path. Needs 2 GPUs."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    from tests.test_sharded_gloo import _free_port
    mp.spawn(_p2p_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r = [np.load(tmp_path / f"p2p{i}.npz") for i in range(2)]
    assert all(bool(x["same"]) and not bool(x["failed"]) for x in r)
    assert (r[0]["ids"] == r[1]["ids"]).all()
