"""Update/delete parity: EdgeScan, pruneDeleteNeighbour, removeInboundEdges and the re-insert of
updated points (vamana.go:223-253, prune.go:12-154, node.go:142-199) on the GPU vs the oracle.

The reference walks toPrune in Go-map order; every pruneDeleteNeighbour writes only its own
node and reads only deleted nodes' edges, so the order is unobservable and the GPU (one CTA
per node) must reproduce the oracle's graph edge for edge. Inserts inside the same call are
compared with the mini-batch schedule set to one point (the 1-worker schedule)."""
import numpy as np
import pytest

from semadb_b200 import synth
from semadb_b200.vamana import (BinaryQuantizerParameters, IndexVamana, IndexVectorChange,
                                IndexVectorVamanaParameters, ProductQuantizerParameters, Quantizer)

pytestmark = pytest.mark.gpu


def _same_graph(oix, g, n_rows):
    adj, odeg = oix.get_graph(n_rows)
    deg, e = g.get_edges(np.arange(1, n_rows, dtype=np.uint64))
    assert (deg == odeg[1:n_rows]).all(), f"degrees differ at rows {np.nonzero(deg != odeg[1:n_rows])[0][:5] + 1}"
    for i in range(n_rows - 1):
        assert (e[i, :deg[i]] == adj[i + 1, :deg[i]]).all(), f"edges of node {i + 1} differ"
    assert g.get_start_overflow().tolist() == oix.start_extra().tolist()


def test_edge_scan_kat_gpu():
    """vamana_test.go:142-175 through the C ABI."""
    g = IndexVamana("scan", IndexVectorVamanaParameters(2), start_vector=np.array([1, 0], np.float32))
    g.set_vectors(np.arange(2, 7, dtype=np.uint64), np.zeros((5, 2), np.float32))
    lists = {2: [3, 6], 3: [2, 4], 4: [3, 5], 5: [4], 6: [2]}
    ids = np.array(sorted(lists), dtype=np.uint64)
    g.set_edges(ids, [len(lists[int(i)]) for i in ids], np.concatenate([lists[int(i)] for i in ids]).astype(np.uint64))
    tp, ts = g.edge_scan([3, 4])
    assert tp.tolist() == [2, 5] and ts.tolist() == [5]


@pytest.mark.parametrize("metric,dim", [("euclidean", 128), ("cosine", 48), ("dot", 100), ("hamming", 256)])
def test_update_delete_reproduces_oracle_graph(metric, dim):
    from tests.helpers import mirror_to_gpu, oracle_graph
    n = 3000
    rng = np.random.Generator(np.random.PCG64(dim))
    if metric == "hamming":
        X = synth.planted_bits(n, dim, seed=3, n_proto=64)
        fresh = synth.planted_bits(700, dim, seed=4, n_proto=64, proto_seed=3)
    else:
        X = synth.latent_gaussian(n, dim, seed=dim, latent=8, normalize=(metric == "cosine"))
        fresh = synth.latent_gaussian(700, dim, seed=dim + 1, w_seed=dim, latent=8, normalize=(metric == "cosine"))
    oix, ids, start = oracle_graph(X, metric, threads=1)
    g = mirror_to_gpu(oix, X, ids, start, metric)
    g.insert_config(min_batch=1, max_batch=1, growth_div=1)
    # one InsertUpdateDelete call: 200 inserts, 300 updates, 200 deletes, 3 absent+nil (skipped)
    new_ids = np.arange(n + 2, n + 202, dtype=np.uint32)
    upd_ids = rng.choice(ids[:1500], size=300, replace=False).astype(np.uint32)
    del_ids = rng.choice(ids[1500:], size=200, replace=False).astype(np.uint32)
    ghost = np.array([9000, 9001, 9002], dtype=np.uint32)
    ch_ids = np.concatenate([new_ids, upd_ids, del_ids, ghost])
    vec = np.concatenate([fresh[:500], np.zeros((203, dim), np.float32)])
    has = np.concatenate([np.ones(500, np.uint8), np.zeros(203, np.uint8)])
    oix.update_delete(ch_ids, vec, has, threads=1)
    g.insert_update_delete_batch(ch_ids.astype(np.uint64), vec, has)
    assert g.count == oix.count
    _same_graph(oix, g, n + 202)
    assert (g.get_vectors(upd_ids) == vec[200:500]).all()
    Q = fresh[500:]
    ref = oix.search(Q, k=10, threads=4)
    gi, gd, gc = g.search_batch(Q, 10, 75)
    assert (gc == ref["counts"]).all() and (gi == ref["ids"].astype(np.uint64)).all()
    assert gd.tobytes() == ref["dists"].tobytes()
    assert not np.isin(gi, del_ids).any()
    # a second round on the already-mutated graph (deletes only)
    del2 = rng.choice(np.setdiff1d(ids, del_ids), size=400, replace=False).astype(np.uint32)
    z = np.zeros((400, dim), np.float32)
    oix.update_delete(del2, z, np.zeros(400, np.uint8), threads=1)
    g.insert_update_delete_batch(del2.astype(np.uint64), z, np.zeros(400, np.uint8))
    _same_graph(oix, g, n + 202)


def test_update_delete_pq_store():
    """DistanceFromPoint = SDC table on a fitted product store (product.go:279-305)."""
    from oracle import oraclelib as O
    n, dim, M, K = 2500, 32, 4, 16
    X = synth.latent_gaussian(n, dim, seed=21, latent=8)
    start = synth.start_vector(dim, 5)
    oix = O.OracleIndex(dim, "euclidean", 75, 64, 1.2, quantizer="product", pq_m=M, pq_k=K, pq_trigger=1000)
    oix.set_start(start)
    ids = np.arange(2, n + 2, dtype=np.uint32)
    oix.insert(ids[:1000], X[:1000], threads=1)
    assert oix.fit(pq_first=0, pq_alias=True) == 1
    oix.insert(ids[1000:], X[1000:], threads=1)
    q = Quantizer("product", product=ProductQuantizerParameters(K, M, 1000))
    g = IndexVamana("pq", IndexVectorVamanaParameters(dim, "euclidean", 75, 64, 1.2, q), start_vector=start)
    g.insert_config(min_batch=1, max_batch=1, growth_div=1)
    g.insert_batch(ids[:1000].astype(np.uint64), X[:1000])
    assert g.fit(0)
    g.insert_batch(ids[1000:].astype(np.uint64), X[1000:])
    _same_graph(oix, g, n + 2)
    rng = np.random.Generator(np.random.PCG64(8))
    upd = rng.choice(ids[:1200], size=100, replace=False).astype(np.uint32)
    dele = rng.choice(ids[1200:], size=150, replace=False).astype(np.uint32)
    ch = np.concatenate([upd, dele])
    vec = np.concatenate([synth.latent_gaussian(100, dim, seed=22, w_seed=21, latent=8), np.zeros((150, dim), np.float32)])
    has = np.concatenate([np.ones(100, np.uint8), np.zeros(150, np.uint8)])
    oix.update_delete(ch, vec, has, threads=1)
    g.insert_update_delete_batch(ch.astype(np.uint64), vec, has)
    _same_graph(oix, g, n + 2)
    assert (g.get_codes(upd) == oix.get_codes(upd)).all()


def test_orphan_overflow_and_start_prune():
    """prune.go:137-151 with the start node already at R edges (see the oracle twin in
    tests/test_oracle_kat.py::test_delete_orphan_goes_to_start_overflow)."""
    from oracle import oraclelib as O
    R = 4
    pts = np.array([[i, 0.5 * i] for i in range(2, 11)], np.float32)
    start = np.array([1, 0], np.float32)
    graph = {1: [2, 3, 4, 5], 2: [6, 3], 3: [2, 4], 4: [3, 5], 5: [4, 8], 6: [7], 7: [6], 8: [9, 10], 9: [8], 10: [8]}
    oix = O.OracleIndex(2, "euclidean", 25, R, 1.2)
    oix.set_start(start)
    oix.set_vectors(np.arange(2, 11, dtype=np.uint32), pts)
    adj = np.full((11, R), 0xFFFFFFFF, dtype=np.uint32)
    deg = np.zeros(11, np.uint16)
    for nid, e in graph.items():
        adj[nid, :len(e)] = e
        deg[nid] = len(e)
    oix.set_graph(adj, deg)
    g = IndexVamana("orphan", IndexVectorVamanaParameters(2, "euclidean", 25, R, 1.2), start_vector=start, relaxed=True)
    g.set_vectors(np.arange(2, 11, dtype=np.uint64), pts)
    ids = np.array(sorted(graph), dtype=np.uint64)
    g.set_edges(ids, [len(graph[int(i)]) for i in ids], np.concatenate([graph[int(i)] for i in ids]).astype(np.uint64))
    tp, ts = g.edge_scan([6])
    assert tp.tolist() == [2, 7] and ts.tolist() == [7]
    z = np.zeros((1, 2), np.float32)
    for victim in (6, 2):
        oix.update_delete(np.array([victim], np.uint32), z, np.zeros(1, np.uint8))
        g.insert_update_delete_batch(np.array([victim], np.uint64), z, np.zeros(1, np.uint8))
        _same_graph(oix, g, 11)
        ref = oix.search(pts, k=3, search_size=25)
        gi, gd, gc = g.search_batch(pts, 3, 25)
        assert (gi == ref["ids"].astype(np.uint64)).all() and gd.tobytes() == ref["dists"].tobytes()
        if victim == 6:
            assert g.get_start_overflow().tolist() == [7]
            assert gi[5, 0] == 7 and gd[5, 0] == 0  # the orphan is found through the overflow edge


def test_mirror_insert_update_delete_and_errors():
    """IndexVamana.InsertUpdateDelete through the host mirror (vamana.go:127-263): reserved
    ids are errors (vamana.go:150-157), self search after update (shard_vector_test.go:408-420),
    graph invariants after a batched (relaxed) mixed call."""
    from semadb_b200._capi import ERR_RESERVED_ID, SdbError
    rng = np.random.Generator(np.random.PCG64(1))
    n, dim = 4000, 16
    X = rng.random((n, dim), dtype=np.float32)
    g = IndexVamana("iud", IndexVectorVamanaParameters(dim), start_seed=7)
    g.insert_update_delete(IndexVectorChange(i + 2, X[i]) for i in range(n))
    for bad in (0, 1):
        with pytest.raises(SdbError) as ei:
            g.insert_update_delete([IndexVectorChange(bad, X[0])])
        assert ei.value.code == ERR_RESERVED_ID
    Y = rng.random((500, dim), dtype=np.float32)
    changes = [IndexVectorChange(i + 2, Y[i]) for i in range(500)]             # updates
    changes += [IndexVectorChange(i + 2, None) for i in range(1000, 1800)]      # deletes
    changes += [IndexVectorChange(n + 2 + i, X[i] + 1) for i in range(300)]     # inserts
    g.insert_update_delete(changes)
    assert g.count == 1 + n - 800 + 300
    alive = np.concatenate([np.arange(2, 1002), np.arange(1802, n + 2), np.arange(n + 2, n + 302)])
    deg, e = g.get_edges(np.concatenate([[1], alive]).astype(np.uint64))
    alive_set = set(alive.tolist()) | {1}
    nodes = [1] + alive.tolist()
    nb = {v: [int(x) for x in e[i, :deg[i]]] for i, v in enumerate(nodes)}
    nb[1] += [int(x) for x in g.get_start_overflow()]
    seen, todo = {1}, [1]
    while todo:
        v = todo.pop()
        for u in nb[v]:
            assert u in alive_set and u != v
            if u not in seen:
                seen.add(u)
                todo.append(u)
    assert seen == alive_set
    ids, d, cnt = g.search_batch(Y, 10, 75)
    assert (ids[:, 0] == np.arange(2, 502)).all() and (d[:, 0] == 0).all()
    assert not np.isin(ids, np.arange(1002, 1802)).any()
