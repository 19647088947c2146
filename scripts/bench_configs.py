#!/usr/bin/env python
"""Per-GPU measurements of the other BASELINE.json configs (bench.py measures C2, the config the
metric is quoted on). One JSON line per config: QPS with queries resident in HBM (CUDA events),
algorithmic bytes per query (n_dist * row_bytes + n_hops * R * 4), achieved
GB/s, recall@10 against the exact GPU flat scan, build time of the batched insert (K8).

  c1   100k x 128 uniform f32 L2, 1k queries (L2-resident: reported, not held against HBM)
  c3   one shard of C3: 1.25M x 384 latent-32, L2-normalised, cosine
  c4   C4-shaped: n x 768 latent-64, dot, PQ M=96 K=256 trained on the first 10k points
  c5b  C5b-shaped: n x 1024 planted-cluster bits, hamming
  c5a  build: 1M x 128 SIFT-shaped, batched insert from empty (points/s)

usage: python scripts/bench_configs.py c1,c3,c4,c5b,c5a [--n-c3 N] [--n-c4 N] [--n-c5b N]
"""
import argparse
import json
import os
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402

from semadb_b200 import synth  # noqa: E402
from semadb_b200.vamana import (IndexVamana, IndexVectorVamanaParameters, ProductQuantizerParameters,  # noqa: E402
                                Quantizer)

L, R, ALPHA, K = 75, 64, 1.2, 10
PEAK = 6650.0
try:
    PEAK = float(json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def time_search(g, Q, steps=10, warmup=3):
    dev = torch.device("cuda", 0)
    B = len(Q)
    d_q = torch.from_numpy(Q).to(dev)
    ids = torch.zeros((B, K), dtype=torch.int64, device=dev)
    d = torch.zeros((B, K), dtype=torch.float32, device=dev)
    c = torch.zeros((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(warmup + 2):
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
        torch.cuda.synchronize()  # the visited-table size adapts between searches that find the stream idle
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(steps):
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if os.environ.get("SDB_PROFILE"):  # ncu --profile-from-start off: capture exactly one search
        torch.cuda.profiler.start()
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    hops, nd = g.last_search_stats(B)
    return ms, hops, nd, ids.cpu().numpy()


def recall(g, Q, got, nq=1000):
    nq = min(nq, len(Q))
    fi, _, _ = g.flat_search_batch(Q[:nq], K)
    return float(np.mean([len(set(got[b].tolist()) & set(fi[b].tolist())) / K for b in range(nq)]))


def report(name, workload, g, Q, row_bytes, build_s, extra=None, want_ids=False):
    ms, hops, nd, got = time_search(g, Q)
    B = len(Q)
    bytes_q = float(nd.mean()) * row_bytes + float(hops.mean()) * R * 4
    gbs = bytes_q * B / (ms * 1e-3) / 1e9
    out = {"config": name, "workload": workload, "qps": B / (ms * 1e-3), "ms_per_batch": ms, "batch": B,
           "mean_hops": float(hops.mean()), "mean_ndist": float(nd.mean()), "bytes_per_query": bytes_q,
           "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / PEAK, "hbm_peak_gbs": PEAK,
           "recall_at_10": recall(g, Q, got), "build_s": build_s}
    if extra:
        out.update(extra)
    print(json.dumps(out), flush=True)
    return got if want_ids else None


def c1():
    X, Q = synth.uniform(100_000, 128, 1), synth.uniform(1000, 128, 2)
    g = IndexVamana("c1", IndexVectorVamanaParameters(128, "euclidean", L, R, ALPHA), start_vector=synth.start_vector(128, 99))
    t = time.time()
    g.insert_batch(np.arange(2, len(X) + 2, dtype=np.uint64), X)
    report("c1", "100k x 128 uniform f32 L2, 1k queries, GPU-built graph (77 MB: L2-resident)", g, Q, 512, time.time() - t,
           {"note": "1k queries < 1776 resident query-warps: the GPU is under-filled; dataset fits in L2"})


def c3(n):
    X = synth.latent_gaussian(n, 384, seed=5, latent=32, normalize=True)
    Q = synth.latent_gaussian(10_000, 384, seed=6, w_seed=5, latent=32, normalize=True)
    g = IndexVamana("c3", IndexVectorVamanaParameters(384, "cosine", L, R, ALPHA), start_vector=synth.start_vector(384, 99))
    t = time.time()
    g.insert_batch(np.arange(2, n + 2, dtype=np.uint64), X)
    report("c3", f"one C3 shard: {n} x 384 latent-32 normalised, cosine, 10k queries, GPU-built graph", g, Q, 1536,
           time.time() - t)


def c4(n):
    dim, M, KC = 768, 96, 256
    X = synth.latent_gaussian(n, dim, seed=8, latent=64)
    Q = synth.latent_gaussian(10_000, dim, seed=9, w_seed=8, latent=64)
    q = Quantizer("product", product=ProductQuantizerParameters(KC, M, 10000))
    g = IndexVamana("c4", IndexVectorVamanaParameters(dim, "dot", L, R, ALPHA, q), start_vector=synth.start_vector(dim, 99))
    ids = np.arange(2, n + 2, dtype=np.uint64)
    t = time.time()
    g.insert_batch(ids[:10000], X[:10000])
    t1 = time.time()
    fitted = g.fit(0)
    t_fit = time.time() - t1
    g.insert_batch(ids[10000:], X[10000:])
    build = time.time() - t
    log(f"c4: fit={fitted} in {t_fit:.2f}s, build {build:.1f}s")
    wl = f"C4-shaped: {n} x 768 latent-64, dot, PQ M=96 K=256 (ADC search, codes 96 B/row), 10k queries"
    a = report("c4", wl, g, Q, M, build,
               {"pq_fit_s": t_fit, "adc_table": "shared memory (bulk-copied once per query, 2 query-warps per SM)",
                "note": "bytes/query counts code rows + adjacency, not the 98 KB/query table copy"}, want_ids=True)
    os.environ["SDB_ADC_GLOBAL"] = "1"
    b = report("c4_global_table", wl, g, Q, M, build, {"adc_table": "global memory, read through L1/L2 (12 query-warps per SM)"},
               want_ids=True)
    os.environ.pop("SDB_ADC_GLOBAL")
    log(f"c4: shared-memory table ids identical to global-table ids: {bool((a == b).all())}")


def c5b(n):
    dim = 1024
    chunk = 250_000
    g = IndexVamana("c5b", IndexVectorVamanaParameters(dim, "hamming", L, R, ALPHA), start_vector=synth.start_vector(dim, 99))
    g.reserve(n + 2)
    t_build = 0.0
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        X = synth.planted_bits(m, dim, seed=7 + s, proto_seed=7)
        t = time.time()
        g.insert_batch(np.arange(2 + s, 2 + s + m, dtype=np.uint64), X)
        t_build += time.time() - t
        log(f"c5b: {s + m} points in, {t_build:.1f}s of insert")
    Q = synth.planted_bits(10_000, dim, seed=1234567, proto_seed=7)
    report("c5b", f"C5b-shaped: {n} x 1024-bit planted clusters, hamming, 10k queries, GPU-built graph", g, Q, dim // 8, t_build)
    for slots in [int(x) for x in os.environ.get("SDB_SWEEP_SLOTS", "").split(",") if x]:
        os.environ["SDB_VT_SLOTS"] = str(slots)
        os.environ["SDB_DEBUG_RETRY"] = "1"
        ms, hops, nd, _ = time_search(g, Q, steps=5, warmup=2)
        os.environ.pop("SDB_DEBUG_RETRY")
        ms, hops, nd, _ = time_search(g, Q, steps=10, warmup=2)
        log(f"c5b sweep: visited slots {slots}: {len(Q) / ms * 1e3:.0f} QPS, {ms:.3f} ms/batch")
    os.environ.pop("SDB_VT_SLOTS", None)


def c5a():
    n = 1_000_000
    X = synth.sift_shaped(n, 128, 3)
    g = IndexVamana("c5a", IndexVectorVamanaParameters(128, "euclidean", L, R, ALPHA), start_vector=synth.start_vector(128, 99))
    if os.environ.get("SDB_INSERT_CONFIG"):  # "min,max,growth_div" (A/B of the mini-batch schedule)
        g.insert_config(*[int(x) for x in os.environ["SDB_INSERT_CONFIG"].split(",")])
    ids = np.arange(2, n + 2, dtype=np.uint64)
    torch.cuda.synchronize()
    t = time.time()
    g.insert_batch(ids, X)
    dt = time.time() - t
    Q = synth.sift_shaped(10_000, 128, 4, w_seed=3)
    ms, hops, nd, got = time_search(g, Q, steps=3)
    deg, _ = g.get_edges(ids[:100000])
    print(json.dumps({"config": "c5a", "workload": "batched build from empty: 1M x 128 SIFT-shaped f32 L2 (host vectors, H2D inside)",
                      "build_s": dt, "points_per_s": n / dt, "recall_at_10_of_built_graph": recall(g, Q, got),
                      "mean_out_degree": float(deg.mean()), "search_qps_on_built_graph": len(Q) / (ms * 1e-3),
                      "insert_config": os.environ.get("SDB_INSERT_CONFIG", "default 1,4096,16")}), flush=True)


def flat(n=1_000_000, B=10_000):
    """K5: IndexFlat.Search for a 10k batch over 1M x 128: tensor-core candidate pass + exact
    re-score vs the exact CUDA-core scan (same results), through the host-buffer C-ABI call."""
    import os
    from semadb_b200.vamana import IndexFlat, IndexVectorFlatParameters
    X = synth.sift_shaped(n, 128, 3)
    Q = synth.sift_shaped(B, 128, 4, w_seed=3)
    g = IndexFlat(IndexVectorFlatParameters(128, "euclidean"))
    g.set_vectors(np.arange(2, n + 2, dtype=np.uint64), X)
    out = {}
    for name, env in (("tensor_core", None), ("exact_cuda_core", "1")):
        if env:
            os.environ["SDB_FLAT_EXACT"] = env
        else:
            os.environ.pop("SDB_FLAT_EXACT", None)
        nq = B if env is None else 2000
        if env and os.environ.get("SDB_FLAT_SKIP_EXACT"):
            continue
        g.flat_search_batch(Q[:nq], K)  # warm-up: scratch allocation, bf16 shadow of the store
        torch.cuda.synchronize()
        reps = 3
        t = time.time()
        for _ in range(reps):
            ids, d, c = g.flat_search_batch(Q[:nq], K)
        dt = (time.time() - t) / reps
        out[name] = {"queries": nq, "seconds": dt, "qps": nq / dt, "tflops_useful": 2.0 * nq * n * 128 / dt / 1e12}
        out[name + "_ids"] = ids
    os.environ.pop("SDB_FLAT_EXACT", None)
    tc_ids = out.pop("tensor_core_ids")
    same = bool((tc_ids[:2000] == out.pop("exact_cuda_core_ids")).all()) if "exact_cuda_core_ids" in out else None
    print(json.dumps({"config": "flat", "workload": f"IndexFlat.Search {B} queries x {n} x 128 f32 L2, k=10 (host buffers in and out)",
                      "identical_ids_first_2000": same, **out}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("configs", nargs="?", default="c1,c3,c4,c5b,c5a")
    ap.add_argument("--n-c3", type=int, default=1_250_000)
    ap.add_argument("--n-c4", type=int, default=1_000_000)
    ap.add_argument("--n-c5b", type=int, default=2_000_000)
    a = ap.parse_args()
    for c in a.configs.split(","):
        t = time.time()
        {"c1": c1, "c3": lambda: c3(a.n_c3), "c4": lambda: c4(a.n_c4), "c5b": lambda: c5b(a.n_c5b), "c5a": c5a, "flat": flat}[c]()
        log(f"{c} done in {time.time() - t:.1f}s")


if __name__ == "__main__":
    main()
