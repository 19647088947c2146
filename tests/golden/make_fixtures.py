#!/usr/bin/env python
"""Regenerates tests/golden/*.npz: frozen outputs of the CPU oracle (oracle/oracle.cpp, itself pinned
by the reference's known-answer tests, see reference_kats.json and tests/test_oracle_kat.py) on
small seeded inputs. The Go reference cannot be built in this image (no Go toolchain), so these
are oracle outputs, not reference outputs; they freeze today's behaviour so that a later change
to either the oracle or the CUDA path shows up as a diff against committed data.

    python tests/golden/make_fixtures.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from oracle import oraclelib as O  # noqa: E402
from semadb_b200 import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


def build(metric, X, start_seed):
    n, dim = X.shape
    ix = O.OracleIndex(dim, metric, 75, 64, 1.2)
    start = synth.start_vector(dim, start_seed)
    ix.set_start(start)
    ids = np.arange(2, n + 2, dtype=np.uint32)
    ix.insert(ids, X, threads=1)  # the reference's sequential (1-worker) schedule
    return ix, ids, start


def main():
    # f32 squared-L2: graph, search, flat
    X = synth.sift_shaped(800, 32, 11)
    Q = synth.sift_shaped(40, 32, 12, w_seed=11)
    ix, ids, start = build("euclidean", X, 5)
    adj, deg = ix.get_graph()
    s = ix.search(Q, k=10, search_size=75, threads=1, diagnostics=True)
    f = ix.flat_search(Q, k=10, threads=1)
    np.savez_compressed(OUT / "l2_800x32.npz", X=X, Q=Q, start=start, adj=adj, deg=deg, ids=s["ids"], dists=s["dists"],
                        counts=s["counts"], hops=s["hops"], ndist=s["ndist"], flat_ids=f["ids"], flat_dists=f["dists"])
    # hamming on 0/1 floats (threshold 0.5, vectorstore.go:56-66)
    Xb = synth.planted_bits(600, 256, seed=21, proto_seed=22)
    Qb = synth.planted_bits(30, 256, seed=23, proto_seed=22)
    ixb, idsb, startb = build("hamming", Xb, 6)
    adjb, degb = ixb.get_graph()
    sb = ixb.search(Qb, k=10, search_size=75, threads=1, diagnostics=True)
    np.savez_compressed(OUT / "hamming_600x256.npz", X=Xb.astype(np.uint8), Q=Qb.astype(np.uint8), start=startb, adj=adjb,
                        deg=degb, ids=sb["ids"], dists=sb["dists"], counts=sb["counts"], hops=sb["hops"], ndist=sb["ndist"])
    print("wrote", sorted(p.name for p in OUT.glob("*.npz")))


if __name__ == "__main__":
    main()
