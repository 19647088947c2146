"""Host mirror of indexManager.searchParallel's merge step (shard/index/search.go:251-298) over
the C ABI: the members of an `_and` / `_or` query have been searched (vector members through
IndexVamana / IndexFlat), and their ranked lists are combined on the GPU by sdb_hybrid_merge —
union or intersection of the result-id sets, HybridScore of duplicates added up, first non-nil
distance kept, sorted by HybridScore descending.

Members that are not ranked searches (inverted-index sets, search.go:137-165) stay in Go."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np

from . import _capi
from ._capi import check, f32p, u32p, u64p
from .vamana import SearchResult


def hybrid_merge(ids, hybrid, dists, counts, disjunction: bool, device: int = 0):
    """Batched form. ids / hybrid / dists: [S, B, k] (sub-search-major; dists NaN = no distance),
    counts: [S, B]. Returns (ids [B, S*k], hybrid, dists, counts [B])."""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    hybrid = np.ascontiguousarray(hybrid, dtype=np.float32)
    dists = np.ascontiguousarray(dists, dtype=np.float32)
    counts = np.ascontiguousarray(counts, dtype=np.uint32)
    S, B, k = ids.shape
    oi = np.zeros((B, S * k), np.uint64)
    oh = np.zeros((B, S * k), np.float32)
    od = np.zeros((B, S * k), np.float32)
    oc = np.zeros(B, np.uint32)
    check(_capi.lib().sdb_hybrid_merge(device, S, B, k, 1 if disjunction else 0, ids.ctypes.data_as(u64p),
                                       hybrid.ctypes.data_as(f32p), dists.ctypes.data_as(f32p), counts.ctypes.data_as(u32p),
                                       oi.ctypes.data_as(u64p), oh.ctypes.data_as(f32p), od.ctypes.data_as(f32p),
                                       oc.ctypes.data_as(u32p)))
    return oi, oh, od, oc


def search_parallel_merge(member_results: Sequence[Sequence[SearchResult]], is_disjunction: bool, device: int = 0):
    """One request, reference-shaped (search.go:251-298): member_results[i] = the SearchResult list
    of member query i (its result set is the ids of that list). Returns (final id set, merged
    results). A single member is returned as is (search.go:246-249)."""
    if len(member_results) == 1:
        res = list(member_results[0])
        return {r.node_id for r in res}, res
    S = len(member_results)
    k = max(1, max(len(r) for r in member_results))
    ids = np.zeros((S, 1, k), np.uint64)
    h = np.zeros((S, 1, k), np.float32)
    d = np.full((S, 1, k), np.nan, np.float32)
    c = np.zeros((S, 1), np.uint32)
    for s, res in enumerate(member_results):
        c[s, 0] = len(res)
        for r, x in enumerate(res):
            ids[s, 0, r] = x.node_id
            h[s, 0, r] = x.hybrid_score
            if x.distance is not None:
                d[s, 0, r] = x.distance
    oi, oh, od, oc = hybrid_merge(ids, h, d, c, is_disjunction, device)
    out: List[SearchResult] = []
    for j in range(int(oc[0])):
        dist = None if np.isnan(od[0, j]) else float(od[0, j])
        out.append(SearchResult(int(oi[0, j]), dist, float(oh[0, j])))
    return {r.node_id for r in out}, out
