"""The distance family through the C ABI (distance/distance.go:11-97) vs the reference's
known-answer tests and the oracle; flat index parity (flat_test.go:134-191); merge (K6)."""
import numpy as np
import pytest

from oracle import oraclelib as O
from semadb_b200 import _capi, synth
from semadb_b200._capi import f32p, u32p, u64p

pytestmark = pytest.mark.gpu


def _dist_float(metric, X, Y):
    X = np.ascontiguousarray(X, np.float32)
    Y = np.ascontiguousarray(Y, np.float32)
    out = np.zeros(len(X), np.float32)
    _capi.check(_capi.lib().sdb_distance_float(_capi.METRICS[metric], 0, len(X), X.shape[1], X.ctypes.data_as(f32p),
                                               Y.ctypes.data_as(f32p), out.ctypes.data_as(f32p)))
    return out


def _dist_bits(metric, X, Y):
    X = np.ascontiguousarray(X, np.uint64)
    Y = np.ascontiguousarray(Y, np.uint64)
    out = np.zeros(len(X), np.float32)
    _capi.check(_capi.lib().sdb_distance_bits(_capi.METRICS[metric], 0, len(X), X.shape[1], X.ctypes.data_as(u64p),
                                              Y.ctypes.data_as(u64p), out.ctypes.data_as(f32p)))
    return out


def test_distance_known_answers():
    # distance/distance_test.go:9-21: dot and squared euclidean, exact
    table = [([0, 0, 0], [0, 0, 0], 0, 0), ([1, 1, 0], [1, 1, 0], 2, 0), ([1, 2, 3], [4, 5, 6], 32, 27),
             ([-1, -2, -3], [-4, -5, -6], 32, 27), ([-1, 2, 3], [4, -5, 6], 4, 83)]
    X = [t[0] for t in table]
    Y = [t[1] for t in table]
    assert _dist_float("euclidean", X, Y).tolist() == [t[3] for t in table]
    assert _dist_float("dot", X, Y).tolist() == [-t[2] for t in table]          # distance.go:19-21
    assert _dist_float("cosine", X, Y).tolist() == [1 - t[2] for t in table]    # distance.go:23-25
    # distance_test.go:41-57
    assert _dist_bits("hamming", [[0b1001, 0b1]], [[0b1101, 0b0]]).tolist() == [2.0]
    assert _dist_bits("jaccard", [[0b1001, 0b1], [0, 0]], [[0b1101, 0b0], [0, 0]]).tolist() == [0.5, 0.0]
    # distance_test.go:59-67
    d = _dist_float("haversine", [[-34.83333, -58.5166646]], [[49.0083899664, 2.53844117956]])[0] / 1000
    assert abs(d - 11099.54) < 0.01


@pytest.mark.parametrize("dim", [1, 2, 7, 31, 32, 33, 100, 128, 384, 768, 1000, 1536, 4096])
def test_float_distances_bit_identical_to_oracle(dim):
    rng = np.random.Generator(np.random.PCG64(dim))
    X = rng.standard_normal((64, dim)).astype(np.float32)
    Y = rng.standard_normal((64, dim)).astype(np.float32)
    for metric in ("euclidean", "dot", "cosine"):
        want = np.array([O.float_dist(metric, X[i], Y[i]) for i in range(64)], np.float32)
        assert _dist_float(metric, X, Y).tobytes() == want.tobytes()


def test_bit_distances_match_oracle():
    rng = np.random.Generator(np.random.PCG64(3))
    X = rng.integers(0, 1 << 63, size=(200, 16), dtype=np.uint64)
    Y = rng.integers(0, 1 << 63, size=(200, 16), dtype=np.uint64)
    for metric in ("hamming", "jaccard"):
        want = np.array([O.bit_dist(metric, X[i], Y[i]) for i in range(200)], np.float32)
        assert _dist_bits(metric, X, Y).tobytes() == want.tobytes()


def test_bq_encode_matches_reference():
    # binary_test.go:11-23
    v = np.array([[1.0, 0.1, 0.6, 0.7, 0.4]], np.float32)
    thr = np.full(5, 0.5, np.float32)
    out = np.zeros((1, 1), np.uint64)
    _capi.check(_capi.lib().sdb_bq_encode(0, 1, 5, v.ctypes.data_as(f32p), thr.ctypes.data_as(f32p),
                                          out.ctypes.data_as(u64p)))
    assert int(out[0, 0]) == 0b01101
    rng = np.random.Generator(np.random.PCG64(9))
    V = rng.random((37, 200), dtype=np.float32)
    T = rng.random(200, dtype=np.float32)
    out = np.zeros((37, 4), np.uint64)
    _capi.check(_capi.lib().sdb_bq_encode(0, 37, 200, V.ctypes.data_as(f32p), T.ctypes.data_as(f32p),
                                          out.ctypes.data_as(u64p)))
    for i in range(37):
        assert (out[i] == O.bq_encode(V[i], T)).all()


@pytest.mark.parametrize("metric", ["euclidean", "cosine", "dot", "haversine"])
def test_flat_index_matches_bruteforce(metric):
    """flat_test.go:134-191: 2000 random 2-d points; distances exactly equal brute force."""
    from semadb_b200.vamana import IndexFlat, IndexVectorChange, IndexVectorFlatParameters, SearchVectorFlatOptions
    rng = np.random.Generator(np.random.PCG64(11))
    X = rng.random((2000, 2), dtype=np.float32)
    if metric == "haversine":
        X = (X * np.float32(90)).astype(np.float32)
    ids = np.arange(2, 2002, dtype=np.uint32)
    f = IndexFlat(IndexVectorFlatParameters(2, metric))
    f.insert_update_delete([IndexVectorChange(int(i), x) for i, x in zip(ids, X)])
    oix = O.OracleIndex(2, metric)
    oix.set_vectors(ids, X)
    gt = oix.flat_search(X[:64], k=10, threads=4)
    fi, fd, fc = f.flat_search_batch(X[:64], 10)
    assert (fc == 10).all()
    if metric == "haversine":
        # float64 sin/cos/asin come from different libms (CUDA vs glibc): 1e-5 relative,
        # the reference's own test allows 0.01 km (distance_test.go:66)
        assert np.allclose(fd, gt["dists"], rtol=1e-5, atol=1.0)
    else:
        assert fd.tobytes() == gt["dists"].tobytes()
        assert (fi == gt["ids"].astype(np.uint64)).all()
    s, res = f.search(SearchVectorFlatOptions(X[3], limit=5))
    assert len(res) == 5 and res[0].hybrid_score == -res[0].distance
    if metric == "euclidean":
        assert res[0].node_id == 5 and res[0].distance == 0
    # filter (flat.go:94-96) and delete (flat.go:52-54)
    fi2, fd2, fc2 = f.flat_search_batch(X[:4], 10, filter_ids=[10, 20, 30])
    assert (fc2 == 3).all() and set(fi2[0, :3].tolist()) == {10, 20, 30}
    f.insert_update_delete([IndexVectorChange(5, None)])
    fi3, _, fc3 = f.flat_search_batch(X[3:4], 75)
    assert 5 not in fi3[0].tolist()


def test_flat_large_dim_and_k75():
    from semadb_b200.vamana import IndexFlat, IndexVectorFlatParameters
    for dim in (768, 1000):
        X = synth.latent_gaussian(3000, dim, seed=dim, latent=8)
        ids = np.arange(2, 3002, dtype=np.uint32)
        f = IndexFlat(IndexVectorFlatParameters(dim, "euclidean"))
        f.set_vectors(ids.astype(np.uint64), X)
        oix = O.OracleIndex(dim, "euclidean")
        oix.set_vectors(ids, X)
        gt = oix.flat_search(X[:40], k=75, threads=8)
        fi, fd, fc = f.flat_search_batch(X[:40], 75)
        assert fd.tobytes() == gt["dists"].tobytes() and (fi == gt["ids"].astype(np.uint64)).all()


def test_merge_topk_matches_oracle():
    """cluster/actions.go:357-376."""
    rng = np.random.Generator(np.random.PCG64(5))
    S, B, k = 8, 300, 10
    d = np.sort(rng.integers(0, 40, size=(S, B, k)).astype(np.float32), axis=2)  # many ties
    ids = rng.integers(2, 1 << 40, size=(S, B, k)).astype(np.uint64)
    cnt = rng.integers(0, k + 1, size=(S, B)).astype(np.uint32)
    oi, od, oc = O.merge_topk(ids, d, cnt, k)
    gi = np.zeros((B, k), np.uint64)
    gd = np.zeros((B, k), np.float32)
    gc = np.zeros(B, np.uint32)
    _capi.check(_capi.lib().sdb_merge_topk(0, S, B, k, ids.ctypes.data_as(u64p), d.ctypes.data_as(f32p),
                                           cnt.ctypes.data_as(u32p), gi.ctypes.data_as(u64p),
                                           gd.ctypes.data_as(f32p), gc.ctypes.data_as(u32p)))
    assert (gc == oc).all() and (gi == oi).all() and gd.tobytes() == od.tobytes()


@pytest.mark.parametrize("disjunction", [True, False])
def test_hybrid_merge_matches_oracle(disjunction):
    """sdb_hybrid_merge vs the oracle's restatement of searchParallel (shard/index/search.go:259-298):
    random overlapping sub-search lists, some without distances, ragged counts."""
    rng = np.random.Generator(np.random.PCG64(77))
    S, B, k = 3, 500, 10
    ids = rng.integers(2, 40, size=(S, B, k)).astype(np.uint64)
    for s in range(S):  # ids are unique inside one sub-search's list
        for b in range(B):
            ids[s, b] = rng.permutation(np.arange(2, 40, dtype=np.uint64))[:k]
    d = np.sort(rng.integers(0, 30, size=(S, B, k)).astype(np.float32), axis=2)
    w = np.array([0.5, 1.0, 2.0], np.float32).reshape(S, 1, 1)
    h = (np.float32(-1) * d * w).astype(np.float32)
    d[2] = np.nan  # a text sub-search: scores, no distances
    h[2] = np.sort(rng.random((B, k)).astype(np.float32), axis=1)[:, ::-1]
    cnt = rng.integers(0, k + 1, size=(S, B)).astype(np.uint32)
    oi, oh, od, oc = O.hybrid_merge(ids, h, d, cnt, disjunction)
    gi = np.zeros((B, S * k), np.uint64)
    gh = np.zeros((B, S * k), np.float32)
    gd = np.zeros((B, S * k), np.float32)
    gc = np.zeros(B, np.uint32)
    _capi.check(_capi.lib().sdb_hybrid_merge(0, S, B, k, 1 if disjunction else 0, ids.ctypes.data_as(u64p),
                                             h.ctypes.data_as(f32p), d.ctypes.data_as(f32p), cnt.ctypes.data_as(u32p),
                                             gi.ctypes.data_as(u64p), gh.ctypes.data_as(f32p), gd.ctypes.data_as(f32p),
                                             gc.ctypes.data_as(u32p)))
    assert (gc == oc).all() and (gi == oi).all()
    assert gh.tobytes() == oh.tobytes() and gd.tobytes() == od.tobytes()
    assert oc.max() > k if disjunction else (oc.max() >= 1 and oc.min() == 0)
