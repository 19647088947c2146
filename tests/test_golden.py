"""Committed golden data (tests/golden/): the reference's own known-answer vectors
(reference_kats.json, transcribed from its Go tests) pin the oracle; the oracle's frozen outputs on
small seeded inputs (*.npz, tests/golden/make_fixtures.py) pin both the oracle against drift (CPU)
and the CUDA path (GPU) — graph edge for edge, ids, distances, hop and distance counters."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import oraclelib as O
from semadb_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
KATS = json.loads((GOLD / "reference_kats.json").read_text())


def test_reference_kats_float_and_bits():
    for c in KATS["float_distance_table"]["cases"]:
        for impl in ("pure", "raw", "model"):
            assert O.float_dist("dot", c["x"], c["y"], impl) == np.float32(c["dot"]), c["name"]
            assert O.float_dist("euclidean", c["x"], c["y"], impl) == np.float32(c["squared_l2"]), c["name"]
    b = KATS["bit_distance"]
    assert O.bit_dist("hamming", b["x"], b["y"]) == b["hamming"] and O.bit_dist("jaccard", b["x"], b["y"]) == b["jaccard"]
    assert O.bit_dist("jaccard", [0, 0], [0, 0]) == b["jaccard_of_zeros"]
    h = KATS["haversine"]
    assert abs(O.float_dist("haversine", h["x"], h["y"]) / 1000 - h["km"]) < h["tolerance_km"]


def test_reference_kats_quantizer_graph_and_cluster():
    e = KATS["binary_encode"]
    assert [int(w) for w in O.bq_encode(e["vector"], e["threshold"])] == e["words"]
    f = KATS["binary_fit"]
    assert O.bq_fit_threshold(np.array(f["vectors"], np.float32)).tolist() == f["threshold"]
    s = KATS["edge_scan"]
    ix = O.OracleIndex(2, "euclidean", 75, 64, 1.2)
    ix.set_start(np.array([1, 0], np.float32))
    n = max(int(k) for k in s["edges"]) + 1
    ix.set_vectors(np.arange(2, n, dtype=np.uint32), np.zeros((n - 2, 2), np.float32))
    adj = np.full((n, 64), 0xFFFFFFFF, dtype=np.uint32)
    deg = np.zeros(n, np.uint16)
    for nid, ed in s["edges"].items():
        adj[int(nid), :len(ed)] = ed
        deg[int(nid)] = len(ed)
    ix.set_graph(adj, deg)
    tp, ts = ix.edge_scan(s["delete"])
    assert tp.tolist() == s["to_prune"] and ts.tolist() == s["to_save"]
    for c in KATS["shard_limit"]["cases"]:
        assert O.shard_limit(c["limit"], c["shards"], c["max"]) == c["want"]


def _oracle_from(fx, metric):
    X = fx["X"].astype(np.float32)
    ix = O.OracleIndex(X.shape[1], metric, 75, 64, 1.2)
    ix.set_start(fx["start"])
    ids = np.arange(2, len(X) + 2, dtype=np.uint32)
    ix.insert(ids, X, threads=1)
    return ix, X, ids


@pytest.mark.parametrize("name,metric", [("l2_800x32", "euclidean"), ("hamming_600x256", "hamming")])
def test_oracle_reproduces_frozen_outputs(name, metric):
    fx = np.load(GOLD / f"{name}.npz")
    ix, X, ids = _oracle_from(fx, metric)
    adj, deg = ix.get_graph()
    assert (deg == fx["deg"]).all() and (adj == fx["adj"]).all()
    s = ix.search(fx["Q"].astype(np.float32), k=10, search_size=75, threads=1, diagnostics=True)
    assert (s["ids"] == fx["ids"]).all() and s["dists"].tobytes() == fx["dists"].tobytes()
    assert (s["hops"] == fx["hops"]).all() and (s["ndist"] == fx["ndist"]).all()
    if "flat_ids" in fx:
        f = ix.flat_search(fx["Q"].astype(np.float32), k=10, threads=1)
        assert (f["ids"] == fx["flat_ids"]).all() and f["dists"].tobytes() == fx["flat_dists"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("name,metric", [("l2_800x32", "euclidean"), ("hamming_600x256", "hamming")])
def test_cuda_path_reproduces_frozen_outputs(name, metric):
    """No oracle in the loop: the CUDA build (mini-batch 1 = the reference's sequential schedule),
    beam search and flat scan against the committed fixture."""
    from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters
    fx = np.load(GOLD / f"{name}.npz")
    X = fx["X"].astype(np.float32)
    Q = fx["Q"].astype(np.float32)
    g = IndexVamana("golden", IndexVectorVamanaParameters(X.shape[1], metric, 75, 64, 1.2), start_vector=fx["start"])
    g.insert_config(1, 1, 1)  # one point per mini-batch
    ids = np.arange(2, len(X) + 2, dtype=np.uint64)
    g.insert_batch(ids, X)
    deg, edges = g.get_edges(np.arange(1, len(X) + 2, dtype=np.uint64))
    assert (deg == fx["deg"][1:]).all()
    R = edges.shape[1]
    want = fx["adj"][1:, :R].astype(np.uint64)
    mask = np.arange(R)[None, :] < deg[:, None]
    assert (edges[mask] == want[mask]).all()
    gi, gd, gc = g.search_batch(Q, 10, 75)
    hops, nd = g.last_search_stats(len(Q))
    assert (gi == fx["ids"].astype(np.uint64)).all() and gd.tobytes() == fx["dists"].tobytes()
    assert (gc == fx["counts"]).all() and (hops == fx["hops"]).all() and (nd == fx["ndist"]).all()
    if "flat_ids" in fx:
        fi, fd, fc = g.flat_search_batch(Q, 10)
        assert (fi == fx["flat_ids"].astype(np.uint64)).all() and fd.tobytes() == fx["flat_dists"].tobytes()
