#!/usr/bin/env python
"""bench.py — Vamana search throughput on the BASELINE.json config (C2) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A "step" is one pass of the hot path over one 10k-query batch: batched greedy beam
search (K1) over a 1M x 128 f32 L2 shard per GPU (R=64, L=75, alpha=1.2, k=10), plus, at
N>1, the NCCL all-gather of per-GPU top-k lists and the merge kernel (K6).

  value   = shard-searches per second over the whole job = n_gpus * batch / step time,
            queries resident in HBM, timed with CUDA events (max over ranks). Every query
            visits every shard (cluster/actions.go:316-376), so the end-user QPS over the
            N x 1M collection is value / n_gpus (reported as config.user_qps).
  e2e     = the same metric through the C-ABI call sdb_search_batch with HOST (pinned)
            query / result buffers: H2D + kernels + D2H inside the timed region.
  roofline= algorithmic bytes per launch (n_dist*512 + n_hops*256 per query, counted by
            the kernel itself) / beam-search kernel time, vs measured HBM copy bandwidth.
  cpu_baseline / --impl reference = the reference-equivalent C++ restatement (oracle/)
            timed on this box's host cores on the same graph. The Go reference itself
            cannot be built in this image (no Go toolchain).

The graph searched is reference-built as BASELINE.json config[1] asks: the oracle's
restatement of insertSinglePoint (concurrent workers, like vamana.go:190-195) builds it
during untimed input preparation; `--graph gpu` uses the CUDA batched insert (K8) instead.
Nothing on the timed GPU path touches oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

DIM, L, R, ALPHA, K = 128, 75, 64, 1.2, 10
_OUT = sys.stdout


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture
    of this same command (profiles/k1_traffic.json); None if absent."""
    p = ROOT / "profiles" / "k1_traffic.json"
    try:
        t = json.loads(p.read_text())
        return float(t["dram_bytes_read"]) + float(t["dram_bytes_write"])
    except Exception:
        return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i",
                 str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_data(n, rank, nq):
    from semadb_b200 import synth
    X = synth.sift_shaped(n, DIM, seed=3 + 1000 * rank, w_seed=3)
    Q = synth.sift_shaped(nq, DIM, seed=4, w_seed=3)
    start = synth.start_vector(DIM, 99 + rank)
    return X, Q, start


def build_oracle_index(X, start, threads):
    from oracle import oraclelib as O
    oix = O.OracleIndex(DIM, "euclidean", L, R, ALPHA)
    oix.set_start(start)
    ids = np.arange(2, len(X) + 2, dtype=np.uint32)
    t = time.time()
    oix.insert(ids, X, threads=threads)
    return oix, ids, time.time() - t


def run_reference(args):
    """--impl reference: the reference-equivalent C++ restatement on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oraclelib as O
    threads = O.hw_threads()
    args.graph = "oracle"
    X, Q, start = make_data(args.n, 0, args.queries)
    oix, ids, tb = build_oracle_index(X, start, threads)
    log(f"[reference] graph built in {tb:.1f}s with {threads} threads")
    nq = min(args.queries, args.ref_queries)
    for _ in range(args.warmup):
        oix.search(Q[:nq], k=K, search_size=L, threads=threads)
    t0 = time.time()
    for _ in range(args.steps):
        oix.search(Q[:nq], k=K, search_size=L, threads=threads)
    dt = (time.time() - t0) / args.steps
    qps = nq / dt
    out = {
        "impl": "reference", "metric": "vamana_search_qps", "value": qps, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": "port",
                         "sample": f"{nq} of {args.queries} queries per step, {threads} threads, one query per thread"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=_OUT, flush=True)


def workload_config(args, world):
    return {"workload": f"C2: SIFT-shaped synthetic {args.n}x{DIM} f32 L2 Vamana shard per GPU (R={R}, L={L}, "
                        f"alpha={ALPHA}), {args.queries}-query batch, k={K}",
            "points_per_gpu": args.n, "dim": DIM, "batch": args.queries, "k": K, "search_size": L,
            "degree_bound": R, "graph": args.graph,
            "l2_policy": "dataset (vectors+adjacency 768 MB/GPU) >> 126 MB L2, no explicit flush",
            "parallelism": f"shard-per-gpu x{world}, queries broadcast, per-GPU top-k exchanged and merged on every GPU" if world > 1
            else "single shard"}


def _quarantine_stdout():
    """Libraries (NCCL's version banner, torchrun notices) write to fd 1; the driver expects
    exactly one JSON line there. Route fd 1 to stderr and keep a private handle for the line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    global _OUT
    _OUT = _quarantine_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=1_000_000, help="points per GPU shard")
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--ref-queries", type=int, default=10_000)
    ap.add_argument("--graph", default="auto", choices=["auto", "oracle", "gpu"])
    ap.add_argument("--exchange", default=os.environ.get("SDB_EXCHANGE", "auto"), choices=["auto", "p2p", "nccl"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--recall-queries", type=int, default=2000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from semadb_b200 import _capi
    from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _capi.lib()

    # ---- input preparation (untimed) -------------------------------------------------
    t0 = time.time()
    X, Q, start = make_data(args.n, rank, args.queries)
    log(f"[rank {rank}] data generated in {time.time() - t0:.1f}s")
    gix = IndexVamana("bench", IndexVectorVamanaParameters(DIM, "euclidean", L, R, ALPHA), device=local_rank,
                      start_vector=start)
    ids = np.arange(2, args.n + 2, dtype=np.uint64)
    oix = None
    ncpu = os.cpu_count() or 1
    if args.graph == "auto":
        # C2 asks for the reference-built graph. Its CPU build needs ~28 s x 16 threads per 1M-point
        # shard; when N ranks share the host cores and fewer than 6 threads are left per rank the
        # shards are built by the CUDA batched insert (K8) instead — same search QPS within 1 %
        # (profiles/r01_ab_k1.txt), and stated in config.graph.
        args.graph = "oracle" if ncpu // world >= 6 else "gpu"
    if args.graph == "oracle":
        oix, _, tb = build_oracle_index(X, start, max(1, ncpu // world))
        log(f"[rank {rank}] reference-built graph: {tb:.1f}s on {max(1, ncpu // world)} threads")
        gix.set_vectors(ids, X)
        adj, deg = oix.get_graph()
        gix.set_graph_dense(adj[1:], deg[1:], first_id=1)
        del adj
    else:
        t0 = time.time()
        gix.insert_batch(ids, X)
        log(f"[rank {rank}] GPU-built graph (K8): {time.time() - t0:.1f}s")

    from semadb_b200.sharded import ShardedSearcher
    B = args.queries
    d_q = torch.from_numpy(Q).to(dev)  # the broadcast query batch, resident on every rank
    searcher = ShardedSearcher(gix, rank, world, exchange=args.exchange)
    stream = torch.cuda.current_stream()
    launches = [0]
    result = [None]

    def step():
        before = gix.launch_count
        result[0] = searcher.search_batch_device(d_q, K, L)  # K1 (+ all-gather + K6 at N>1)
        # + K6 merge, + the peer barrier kernel when the exchange is fused (NCCL's kernels are not ours)
        launches[0] += gix.launch_count - before + (0 if world == 1 else 2 if searcher.exchange != "nccl" else 1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    n_launch = launches[0]

    # beam-search kernel alone (the dominant kernel), same stream, same inputs
    ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k_ids = torch.zeros((B, K), dtype=torch.int64, device=dev)
    k_d = torch.zeros((B, K), dtype=torch.float32, device=dev)
    k_c = torch.zeros((B,), dtype=torch.int32, device=dev)
    kern_ms = []
    for _ in range(min(args.steps, 10)):
        torch.cuda.synchronize()
        ks.record(stream)
        gix.search_batch_device(d_q, K, L, k_ids, k_d, k_c, stream.cuda_stream)
        ke.record(stream)
        torch.cuda.synchronize()
        kern_ms.append(ks.elapsed_time(ke))
    kern_ms = float(np.mean(kern_ms))
    hops, ndist = gix.last_search_stats(B)

    # ---- end-to-end through the C-ABI with host buffers --------------------------------
    h_q = torch.from_numpy(Q).pin_memory()
    h_ids = torch.zeros((B, K), dtype=torch.int64).pin_memory()
    h_d = torch.zeros((B, K), dtype=torch.float32).pin_memory()
    h_c = torch.zeros((B,), dtype=torch.int32).pin_memory()
    import ctypes as C

    def e2e_step():
        if world == 1:
            _capi.check(lib.sdb_search_batch(gix._h, B, C.cast(h_q.data_ptr(), _capi.f32p), K, L, None, 0,
                                             C.cast(h_ids.data_ptr(), _capi.u64p), C.cast(h_d.data_ptr(), _capi.f32p),
                                             C.cast(h_c.data_ptr(), _capi.u32p)))
            return
        # N > 1: every rank receives the broadcast query batch in host memory (the Go cluster layer
        # fans requests out to shards, cluster/actions.go:316-351), copies it in, runs the sharded
        # search (K1 with the fused peer gather, barrier, K6) and receives the merged lists.
        # The buffers are page-locked: the kernels read / write them in place (mapped host memory).
        searcher.search_batch_pinned(h_q, K, L, h_ids, h_d, h_c, dev)

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop()

    if world > 1:
        t = torch.tensor([ms_total, e2e_s, kern_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, e2e_s, kern_ms = [float(x) for x in t.tolist()]

    # ---- quality + cpu baseline (untimed, rank 0) --------------------------------------
    recall = None
    cpu = None
    parity = None
    if rank == 0:
        nq = min(args.recall_queries, B)
        fi, fd, fc = gix.flat_search_batch(Q[:nq], K)  # exact ground truth on the GPU (K5)
        got = k_ids.cpu().numpy()[:nq]  # this rank's shard-local result
        recall = float(np.mean([len(set(got[b].tolist()) & set(fi[b].tolist())) / K for b in range(nq)]))
        if not args.no_cpu_baseline and world == 1:
            from oracle import oraclelib as O
            threads = O.hw_threads()
            if oix is None:
                oix, _, _ = build_oracle_index(X, start, threads)
            oix.search(Q[:1000], k=K, search_size=L, threads=threads)
            t0 = time.time()
            reps = 0
            while reps < 3 or (time.time() - t0 < 5.0 and reps < 50):
                ref = oix.search(Q, k=K, search_size=L, threads=threads)
                reps += 1
            cdt = (time.time() - t0) / reps
            t1 = time.time()
            oix.search(Q[:2000], k=K, search_size=L, threads=1)
            qps1 = 2000 / (time.time() - t1)
            cpu = {"value": B / cdt, "unit": "queries/s", "cores": threads, "kind": "port",
                   "sample": f"all {B} queries x {reps} passes, {threads} threads, one query per thread",
                   "single_thread_qps": qps1}
            if args.graph == "oracle":
                same = (k_ids.cpu().numpy() == ref["ids"].astype(np.int64)).all(axis=1).mean()
                parity = {"id_rows_identical_to_oracle": float(same),
                          "dists_bit_identical": bool(k_d.cpu().numpy().tobytes() == ref["dists"].tobytes())}

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * B / (ms_step * 1e-3)
        bytes_q = float(ndist.mean()) * DIM * 4 + float(hops.mean()) * R * 4
        achieved = bytes_q * B / (kern_ms * 1e-3) / 1e9
        peak, peak_kind = measured_peaks()
        cfg = workload_config(args, world)
        if world > 1:
            cfg.update(exchange="fused peer stores over NVLink + flag barrier (sdb_search_batch_gather_device)"
                       if searcher.exchange != "nccl" else "NCCL all-gather per result tensor",
                       peer_barrier_timed_out=(searcher._peer.barrier_failed() if searcher._peer is not None else None))
        cfg.update(user_qps=B / (ms_step * 1e-3), recall_at_10=recall, mean_hops=float(hops.mean()),
                   mean_ndist=float(ndist.mean()), bytes_per_query=bytes_q, parity=parity,
                   host_cores=os.cpu_count())
        out = {
            "metric": "vamana_search_qps", "value": value, "unit": "queries/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak,
                         "traffic": measured_traffic() if (args.n == 1_000_000 and B == 10_000) else None,
                         "algorithmic_bytes_per_launch": bytes_q * B, "peak_kind": peak_kind,
                         "kernel": "beam_search_kernel", "kernel_ms": kern_ms},
            "cpu_baseline": cpu,
            "e2e": {"value": world * B / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": B * DIM * 4,
                    "d2h_bytes_per_step": B * K * 12 + B * 4, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": n_launch,
            "clocks": clocks,
        }
        print(json.dumps(out), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
