"""CPU-side checks of the drop-in boundary: the library loads and exports every symbol
include/semadb_b200.h declares; no compute is attempted without a GPU."""
import ctypes

import pytest

from semadb_b200 import _capi


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = _capi.declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/semadb_b200.h but not exported"
    assert set(names) == set(_capi._SIGS), "ctypes table and header disagree"


def test_abi_version_and_pure_host_functions():
    L = _capi.lib()
    assert L.sdb_abi_version() == 1
    # cluster/actions.go:291-299
    assert L.sdb_shard_limit(10, 8, 75) == 10
    assert L.sdb_shard_limit(100, 5, 75) == 38
    assert L.sdb_shard_limit(75, 1, 75) == 75


def test_no_cpu_fallback():
    """Without a CUDA device index creation must fail loudly (never fall back)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters
    with pytest.raises(_capi.SdbError) as ei:
        IndexVamana("x", IndexVectorVamanaParameters(8), start_seed=1)
    assert ei.value.code == _capi.ERR_CUDA


def test_parameter_validation_mirrors_reference():
    """models/index.go:284-313 ranges are enforced before any CUDA work."""
    from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters, Quantizer, ProductQuantizerParameters
    bad = [
        IndexVectorVamanaParameters(0),
        IndexVectorVamanaParameters(5000),
        IndexVectorVamanaParameters(8, search_size=10),
        IndexVectorVamanaParameters(8, search_size=80),
        IndexVectorVamanaParameters(8, degree_bound=16),
        IndexVectorVamanaParameters(8, degree_bound=100),
        IndexVectorVamanaParameters(8, alpha=1.0),
        IndexVectorVamanaParameters(8, alpha=2.0),
        IndexVectorVamanaParameters(3, "haversine"),
        IndexVectorVamanaParameters(10, quantizer=Quantizer("product", product=ProductQuantizerParameters(256, 3, 1000))),
        IndexVectorVamanaParameters(8, quantizer=Quantizer("product", product=ProductQuantizerParameters(300, 2, 1000))),
    ]
    for p in bad:
        with pytest.raises(_capi.SdbError) as ei:
            IndexVamana("x", p, start_seed=1)
        assert ei.value.code == _capi.ERR_INVALID, p
    with pytest.raises(_capi.SdbError):
        IndexVamana("x", IndexVectorVamanaParameters(8, "manhattan"), start_seed=1)


def test_search_parallel_merge_single_member_passthrough():
    """search.go:246-249: one member query is returned as is — no device call involved."""
    from semadb_b200.search import search_parallel_merge
    from semadb_b200.vamana import SearchResult
    res = [SearchResult(7, 1.5, -1.5), SearchResult(9, 2.0, -2.0)]
    s, out = search_parallel_merge([res], is_disjunction=True)
    assert s == {7, 9} and out == res
