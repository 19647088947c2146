"""The reference's hybrid tests restated over the GPU path (shard/index/search_test.go:411-457,
TestSearch_OrVector): the same five neighbours found by a vamana search and a flat search with
weight 0.5 each merge into five results whose HybridScore is -distance; `_and` of two searches
keeps only the common ids."""
import numpy as np
import pytest

from semadb_b200.search import search_parallel_merge
from semadb_b200.vamana import (IndexFlat, IndexVamana, IndexVectorFlatParameters, IndexVectorVamanaParameters,
                                SearchVectorFlatOptions, SearchVectorVamanaOptions)

pytestmark = pytest.mark.gpu


def _points(n=100):
    # populateIndex (search_test.go): point i has vector [i, i+1], node ids from 2
    return np.arange(2, n + 2, dtype=np.uint64), np.array([[i, i + 1] for i in range(n)], dtype=np.float32)


def test_or_vector_like_the_reference():
    ids, X = _points()
    v = IndexVamana("vector", IndexVectorVamanaParameters(2, "euclidean", 75, 64, 1.2), start_seed=3)
    v.insert_batch(ids, X)
    f = IndexFlat(IndexVectorFlatParameters(2, "euclidean"))
    f.set_vectors(ids, X)
    q = [42.0, 43.0]
    _, rv = v.search(SearchVectorVamanaOptions(q, 75, 5, 0.5))
    _, rf = f.search(SearchVectorFlatOptions(q, 5, 0.5))
    want = {int(ids[i]) for i in (40, 41, 42, 43, 44)}
    rset, res = search_parallel_merge([rv, rf], is_disjunction=True)
    assert rset == want and len(res) == 5
    assert res[0].node_id == int(ids[42])
    for a, b in zip(res, res[1:]):
        assert a.hybrid_score >= b.hybrid_score
    for r in res:  # two weights of 0.5 on the same distance add up to -distance
        assert r.hybrid_score == -r.distance
    # `_and` of the 5 nearest with the 3 nearest keeps the 3
    _, rv3 = v.search(SearchVectorVamanaOptions(q, 75, 3, 1.0))
    rset, res = search_parallel_merge([rf, rv3], is_disjunction=False)
    assert rset == {int(ids[i]) for i in (41, 42, 43)} and [r.node_id for r in res][0] == int(ids[42])
    # a single member is passed through (search.go:246-249)
    rset, res = search_parallel_merge([rf], is_disjunction=False)
    assert res == rf
