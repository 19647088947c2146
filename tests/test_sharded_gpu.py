"""Sharded search on the GPU: S independent shards searched by K1, lists merged by K6,
compared with the oracle doing the same (cluster/actions.go:316-376)."""
import numpy as np
import pytest

from oracle import oraclelib as O
from semadb_b200 import _capi, sharded, synth

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
