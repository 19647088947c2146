#!/usr/bin/env python
"""Round-2 data experiments at shard size on one GPU: which synthetic generators give a graph the
reference's parameters (L=75, R=64, alpha=1.2) can search at recall@10 >= 0.95 (C3, C5b) and a PQ
search that means something (C4). One JSON line per case. Recall is strict and tie-aware
(hit if dist <= k-th true distance)."""
import json, sys, time
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from semadb_b200 import synth
from semadb_b200.vamana import IndexVamana, IndexVectorVamanaParameters, ProductQuantizerParameters, Quantizer
L, R, ALPHA, K = 75, 64, 1.2, 10
dev = torch.device("cuda", 0)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def search_and_recall(g, Q, nq=1000):
    B = len(Q)
    d_q = torch.from_numpy(Q).to(dev)
    ids = torch.zeros((B, K), dtype=torch.int64, device=dev)
    d = torch.zeros((B, K), dtype=torch.float32, device=dev)
    c = torch.zeros((B,), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream()
    for _ in range(4):
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5):
        g.search_batch_device(d_q, K, L, ids, d, c, st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    hops, nd = g.last_search_stats(B)
    got, gd = ids.cpu().numpy()[:nq], d.cpu().numpy()[:nq]
    fi, fd, _ = g.flat_search_batch(Q[:nq], K)
    strict = float(np.mean([len(set(got[b].tolist()) & set(fi[b].tolist())) / K for b in range(nq)]))
    tie = float(np.mean([(gd[b] <= fd[b, K - 1]).sum() / K for b in range(nq)]))
    return {"qps": B / ms * 1e3, "ms": ms, "hops": float(hops.mean()), "ndist": float(nd.mean()),
            "recall_strict": strict, "recall_tie_aware": tie}


def build_chunks(g, gen, n):
    g.reserve(n + 2)
    t_ins = 0.0
    for s, x in gen:
        xh = x.cpu().numpy()
        t = time.time()
        g.insert_batch(np.arange(2 + s, 2 + s + len(xh), dtype=np.uint64), xh)
        t_ins += time.time() - t
    return t_ins


def c3(n):
    g = IndexVamana("c3", IndexVectorVamanaParameters(384, "cosine", L, R, ALPHA), start_vector=synth.start_vector(384, 99))
    tb = build_chunks(g, synth.latent_gaussian_torch(n, 384, 5, dev, latent=16, normalize=True), n)
    Q = torch.cat([x for _, x in synth.latent_gaussian_torch(10000, 384, 6, dev, w_seed=5, latent=16, normalize=True)]).cpu().numpy()
    print(json.dumps({"case": "c3_latent16_norm", "n": n, "build_s": tb, **search_and_recall(g, Q)}), flush=True)


def c5b(n, kind):
    g = IndexVamana("c5b", IndexVectorVamanaParameters(1024, "hamming", L, R, ALPHA), start_vector=synth.start_vector(1024, 99))
    if kind == "sign":
        tb = build_chunks(g, synth.sign_bits_torch(n, 1024, 7, dev, latent=16), n)
        Q = torch.cat([x for _, x in synth.sign_bits_torch(10000, 1024, 10, dev, w_seed=7, latent=16)]).cpu().numpy()
    else:
        def gen():
            for s in range(0, n, 250000):
                m = min(250000, n - s)
                yield s, torch.from_numpy(synth.planted_bits(m, 1024, seed=7 + s, proto_seed=7))
        tb = build_chunks(g, gen(), n)
        Q = synth.planted_bits(10000, 1024, seed=1234567, proto_seed=7)
    print(json.dumps({"case": f"c5b_{kind}", "n": n, "build_s": tb, **search_and_recall(g, Q)}), flush=True)


def c4(n, centres):
    dim, M, KC = 768, 96, 256
    q = Quantizer("product", product=ProductQuantizerParameters(KC, M, 10000))
    g = IndexVamana("c4", IndexVectorVamanaParameters(dim, "dot", L, R, ALPHA, q), start_vector=synth.start_vector(dim, 99))
    g.reserve(n + 2)
    t_ins, t_fit, fitted = 0.0, 0.0, False
    kw = dict(latent=16, normalize=True, centres=centres)
    for s, x in synth.latent_gaussian_torch(n, dim, 8, dev, chunk=1 << 17, **kw):
        xh = x.cpu().numpy()
        ids = np.arange(2 + s, 2 + s + len(xh), dtype=np.uint64)
        t = time.time()
        if not fitted:
            g.insert_batch(ids[:10000], xh[:10000])
            t1 = time.time()
            fitted = g.fit(0)
            t_fit = time.time() - t1
            g.insert_batch(ids[10000:], xh[10000:])
        else:
            g.insert_batch(ids, xh)
        t_ins += time.time() - t
    Q = torch.cat([x for _, x in synth.latent_gaussian_torch(10000, dim, 9, dev, w_seed=8, **kw)]).cpu().numpy()
    print(json.dumps({"case": f"c4_clustered{centres}", "n": n, "build_s": t_ins, "fit_s": t_fit, **search_and_recall(g, Q)}), flush=True)


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        name, *rest = spec.split(":")
        t = time.time()
        if name == "c3":
            c3(int(rest[0]))
        elif name == "c5b":
            c5b(int(rest[0]), rest[1])
        elif name == "c4":
            c4(int(rest[0]), int(rest[1]))
        log(f"{spec} done in {time.time() - t:.1f}s")
