"""Shared helpers for parity tests: build a graph with the oracle, mirror it onto the GPU."""
from __future__ import annotations

import numpy as np

from oracle import oraclelib as O
from semadb_b200 import synth
from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters, Quantizer


def oracle_graph(X, metric="euclidean", L=75, R=64, alpha=1.2, start_seed=99, threads=8, **okw):
    """Oracle index with ids 2..n+1 inserted (sequential if threads == 1)."""
    n, dim = X.shape
    ix = O.OracleIndex(dim, metric, L, R, alpha, **okw)
    start = synth.start_vector(dim, start_seed)
    ix.set_start(start)
    ids = np.arange(2, n + 2, dtype=np.uint32)
    ix.insert(ids, X, threads=threads)
    return ix, ids, start


def mirror_to_gpu(oix, X, ids, start, metric="euclidean", L=75, R=64, alpha=1.2, quantizer=None, relaxed=False):
    """GPU IndexVamana holding the same vectors and the oracle-built edges."""
    params = IndexVectorVamanaParameters(X.shape[1], metric, L, R, alpha, quantizer)
    g = IndexVamana("parity", params, start_vector=start, relaxed=relaxed)
    g.set_vectors(ids.astype(np.uint64), X)
    adj, deg = oix.get_graph()
    g.set_graph_dense(adj[1:], deg[1:], first_id=1)
    return g


def recall_at_k(ids, gt_ids):
    B, k = gt_ids.shape
    return float(np.mean([len(set(ids[b].tolist()) & set(gt_ids[b].tolist())) / k for b in range(B)]))
