"""Synthetic inputs for the BASELINE.json configs.

All generators are numpy.random.Generator(PCG64(seed)), float32, deterministic.
Shared by tests/ and bench.py; no oracle or CUDA dependency.
"""
from __future__ import annotations

import numpy as np


def uniform(n: int, dim: int, seed: int) -> np.ndarray:
    """C1: iid U[0,1) like internal/loadrand/loadrand.go:17-23."""
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.random((n, dim), dtype=np.float32)


def _latent_w(latent: int, dim: int, seed: int) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((latent, dim), dtype=np.float32)


def _latent_points(n: int, latent: int, dim: int, w: np.ndarray, seed: int, chunk: int = 1 << 18) -> np.ndarray:
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    out = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        z = rng.standard_normal((m, latent), dtype=np.float32)
        eps = rng.standard_normal((m, dim), dtype=np.float32)
        out[s:s + m] = z @ w + np.float32(0.1) * eps
    return out


def sift_shaped(n: int, dim: int = 128, seed: int = 3, w_seed: int | None = None, latent: int = 16) -> np.ndarray:
    """C2 "SIFT-shaped": z~N(0,I_16), x = clip(round(32*(zW+0.1e)+64), 0, 255).
    Queries use their own seed but the data's W (w_seed = data seed)."""
    w = _latent_w(latent, dim, seed if w_seed is None else w_seed)
    x = _latent_points(n, latent, dim, w, seed)
    np.multiply(x, np.float32(32), out=x)
    np.add(x, np.float32(64), out=x)
    np.rint(x, out=x)
    np.clip(x, 0, 255, out=x)
    return x


def latent_gaussian(n: int, dim: int, seed: int, w_seed: int | None = None, latent: int = 16,
                    normalize: bool = False) -> np.ndarray:
    """C3 (latent-16, L2-normalised, cosine) / C4 (latent-64, raw, dot)."""
    w = _latent_w(latent, dim, seed if w_seed is None else w_seed)
    x = _latent_points(n, latent, dim, w, seed)
    if normalize:
        nrm = np.sqrt((x.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
        x /= np.maximum(nrm, np.float32(1e-30))
    return x


def planted_bits(n: int, dim: int = 1024, seed: int = 7, n_proto: int = 65536, flip: float = 0.1,
                 proto_seed: int | None = None, chunk: int = 1 << 16) -> np.ndarray:
    """C5b: prototypes Bernoulli(0.5); each point = a random prototype with `flip` of its
    bits flipped; returned as 0.0/1.0 float32 (binarised with threshold 0.5,
    shard/vectorstore/vectorstore.go:56-66)."""
    prng = np.random.Generator(np.random.PCG64(seed if proto_seed is None else proto_seed))
    protos = prng.random((n_proto, dim), dtype=np.float32) < 0.5
    rng = np.random.Generator(np.random.PCG64(seed + 104729))
    out = np.empty((n, dim), dtype=np.float32)
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        pick = rng.integers(0, n_proto, size=m)
        flips = rng.random((m, dim), dtype=np.float32) < flip
        out[s:s + m] = np.logical_xor(protos[pick], flips)
    return out


def start_vector(dim: int, seed: int) -> np.ndarray:
    """setupStartNode (shard/index/vamana/vamana.go:100-110): U(-1,1)^dim, L2-normalised,
    sum accumulated in f32 in index order like the reference loop."""
    rng = np.random.Generator(np.random.PCG64(seed))
    v = (rng.random(dim, dtype=np.float32) * np.float32(2) - np.float32(1)).astype(np.float32)
    s = np.float32(0)
    for x in v:
        s = np.float32(s + np.float32(x * x))
    norm = np.float32(1) / np.float32(np.sqrt(np.float64(s)))
    return (v * norm).astype(np.float32)
