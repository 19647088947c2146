#!/usr/bin/env python
"""A/B the dim-128 beam-search kernel variants (SDB_K1_VARIANT) on one GPU-built C2 graph:
per variant QPS (CUDA events) and equality of ids/dists with variant 0.
usage: python scripts/ab_k1.py [n] [variants,comma,separated]"""
import os, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from semadb_b200 import synth
from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
variants = (sys.argv[2] if len(sys.argv) > 2 else "6:1,6,0,3,7").split(",")  # variant[:flags]
B, K, L = 10_000, 10, 75
X = synth.sift_shaped(n, 128, 3)
Q = synth.sift_shaped(B, 128, 4, w_seed=3)
start = synth.start_vector(128, 99)
g = IndexVamana("ab", IndexVectorVamanaParameters(128, "euclidean", 75, 64, 1.2), device=0, start_vector=start)
t0 = time.time()
g.insert_batch(np.arange(2, n + 2, dtype=np.uint64), X)
print(f"graph built in {time.time()-t0:.1f}s", flush=True)
dev = torch.device("cuda", 0)
d_q = torch.from_numpy(Q).to(dev)
ids = torch.zeros((B, K), dtype=torch.int64, device=dev)
d = torch.zeros((B, K), dtype=torch.float32, device=dev)
c = torch.zeros((B,), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream()
base = None
for v in variants:
    os.environ["SDB_K1_VARIANT"] = v.split(":")[0]
    os.environ["SDB_K1_FLAGS"] = v.split(":")[1] if ":" in v else "0"
    for _ in range(3):
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(10):
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    hops, nd = g.last_search_stats(B)
    out = (ids.cpu().numpy().copy(), d.cpu().numpy().copy(), hops.copy(), nd.copy())
    if base is None:
        base = out
    same = all((a == b).all() for a, b in zip(out, base))
    bytes_q = nd.mean() * 512 + hops.mean() * 256
    print(f"variant {v}: {ms:.3f} ms  {B/ms/1e3:.3f} MQPS  {bytes_q*B/ms/1e6:.0f} GB/s  identical_to_first={same}", flush=True)
