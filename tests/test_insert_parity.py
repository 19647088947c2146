"""K8 parity: the CUDA batched insert (greedySearch + robustPrune + back-edges).

With a mini-batch of one point the GPU build follows the reference's sequential schedule
(the 1-worker case of vamana.go:190-195) and must reproduce the oracle's graph edge for
edge. Larger mini-batches are judged like the reference's own concurrent insert: by graph
invariants (vamana_test.go:29-75) and recall against the oracle-built graph."""
import numpy as np
import pytest

from semadb_b200 import synth
from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters

pytestmark = pytest.mark.gpu


def _gpu_build(X, start, metric="euclidean", max_batch=None, relaxed=False, R=64, L=75, alpha=1.2):
    g = IndexVamana("build", IndexVectorVamanaParameters(X.shape[1], metric, L, R, alpha), start_vector=start,
                    relaxed=relaxed)
    if max_batch is not None:
        g.insert_config(min_batch=1, max_batch=max_batch, growth_div=16)
    ids = np.arange(2, len(X) + 2, dtype=np.uint64)
    g.insert_batch(ids, X)
    return g, ids


def _edges(g, n):
    deg, e = g.get_edges(np.arange(1, n + 2, dtype=np.uint64))
    return deg, e


@pytest.mark.parametrize("metric,dim,n", [("euclidean", 128, 1500), ("euclidean", 2, 800), ("dot", 96, 800),
                                          ("cosine", 100, 800)])
def test_sequential_insert_reproduces_oracle_graph(metric, dim, n):
    from tests.helpers import oracle_graph
    X = synth.latent_gaussian(n, dim, seed=dim + 5, latent=min(8, dim), normalize=(metric == "cosine"))
    oix, _, start = oracle_graph(X, metric, threads=1)
    g, _ = _gpu_build(X, start, metric, max_batch=1)
    adj, odeg = oix.get_graph()
    deg, e = _edges(g, n)
    assert (deg == odeg[1:n + 2]).all()
    for r in range(n + 1):
        assert e[r, :deg[r]].tolist() == adj[r + 1, :odeg[r + 1]].tolist(), f"node {r + 1}"


def _bfs(deg, e):
    seen = {1}
    stack = [1]
    while stack:
        u = stack.pop()
        for v in e[u - 1, :deg[u - 1]]:
            v = int(v)
            if v not in seen:
                seen.add(v)
                stack.append(v)
    return seen


@pytest.mark.parametrize("n", [1, 100, 4242])
def test_batched_insert_connectivity(n):
    """vamana_test.go:29-46,63-75: BFS from the start node reaches every point."""
    X = synth.uniform(n, 2, 17)
    g, ids = _gpu_build(X, synth.start_vector(2, 3))
    deg, e = _edges(g, n)
    assert deg.max() <= 64
    assert len(_bfs(deg, e)) == n + 1
    for r in range(n + 1):
        row = e[r, :deg[r]].tolist()
        assert (r + 1) not in row and len(set(row)) == len(row)


def test_insert_rejects_reserved_ids():
    """vamana_test.go:77-90."""
    from semadb_b200._capi import ERR_RESERVED_ID, SdbError
    g = IndexVamana("r", IndexVectorVamanaParameters(2), start_seed=1)
    for bad in (0, 1):
        with pytest.raises(SdbError) as ei:
            g.insert_batch(np.array([bad], dtype=np.uint64), np.zeros((1, 2), np.float32))
        assert ei.value.code == ERR_RESERVED_ID


def test_batched_build_quality_matches_oracle_graph():
    """Graph built by the GPU in growing mini-batches: recall@10 within 0.005 of searches
    on the oracle-built graph (SURVEY.md §7.3-⑦), self-recall like vamana_test.go:230-252."""
    from tests.helpers import oracle_graph, recall_at_k
    n = 50_000
    X = synth.sift_shaped(n, 128, 3)
    Q = synth.sift_shaped(1000, 128, 4, w_seed=3)
    oix, _, start = oracle_graph(X)
    g, ids = _gpu_build(X, start)
    gt = oix.flat_search(Q, k=10, threads=8)
    ref = oix.search(Q, k=10, threads=8)
    gi, gd, gc = g.search_batch(Q, 10, 75)
    r_ref = recall_at_k(ref["ids"], gt["ids"])
    r_gpu = recall_at_k(gi, gt["ids"].astype(np.uint64))
    assert r_gpu >= r_ref - 0.005, (r_gpu, r_ref)
    si, sd, sc = g.search_batch(X[:500], 10, 75)
    assert (si[:, 0] == ids[:500]).mean() >= 0.99
    deg, e = _edges(g, n)
    assert len(_bfs(deg, e)) == n + 1
