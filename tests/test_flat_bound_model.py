"""CPU model of the candidate bound of the two-pass flat search (csrc/flat_tc.cu: ext_min_kernel,
tc5_filter_kernel in minimum mode, kth_thresh_kernel, the sign filter): with bf16 operands
emulated in numpy, the candidate set it keeps must contain the exact top-k of every query —
the property the CUDA path's bit-exactness rests on (flat.go:76-132 returns the exact top-k).
The constants are the kernel's (T5_C1, T5_UP, c2); the GPU tests check the CUDA code itself
against the exact scan (tests/test_flat_tc.py)."""
import numpy as np
import pytest

from semadb_b200 import synth

T5_C1 = 0.0157       # flat_tc.cu: 2 (2u + u^2), u = 2^-8
T5_UP = 1.00001
TILE = 256


def bf16_rn(a):
    """float32 -> nearest bf16 (ties to even), returned as float32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def bf16_ru(a):
    """float32 >= 0 -> smallest bf16 >= a."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    down = (a.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    up = (down.view(np.uint32) + np.uint32(0x10000)).view(np.float32)
    return np.where(down == a, down, up).astype(np.float32)


def candidates_two_pass(X, Q, k, metric, sample_div=8):
    n, dim = X.shape
    l2 = metric == "euclidean"
    mu = X[:: max(1, n // 65536)].mean(axis=0, dtype=np.float32) if l2 else np.zeros(dim, np.float32)
    Xc, Qc = (X - mu).astype(np.float32), (Q - mu).astype(np.float32)
    xn, qn = (Xc * Xc).sum(1, dtype=np.float32), (Qc * Qc).sum(1, dtype=np.float32)
    scale = np.float32(-2.0 if l2 else -1.0)
    acc = bf16_rn(scale * Qc) @ bf16_rn(Xc).T          # fp32 accumulate (another order than the MMA's: inside c2)
    a = acc + (xn[None, :] if l2 else 0.0)              # approximate score: d - |q|^2 (L2), -dot (dot), d - 1 (cosine)
    ce = np.float32(T5_C1 if l2 else 0.5 * T5_C1)
    e_up = bf16_ru(ce * np.sqrt(qn) * np.float32(T5_UP))[:, None] * bf16_ru(np.sqrt(xn) * np.float32(T5_UP))[None, :]
    # minimum mode over every stride-th whole tile, groups = (tile, column half)
    whole = n // TILE
    tiles_a = min(whole, max(whole // sample_div, 2 * k, 32))
    stride = whole // tiles_a
    A = a + e_up
    mins = []
    for t in range(tiles_a):
        p0 = t * stride * TILE
        mins.append(A[:, p0:p0 + 128].min(axis=1))
        mins.append(A[:, p0 + 128:p0 + 256].min(axis=1))
    mk = np.sort(np.stack(mins, axis=1), axis=1)[:, k - 1]
    x2 = xn.max()
    c2 = np.float32((dim + 32) * 4.8e-7)
    F = c2 * (qn + x2) if l2 else c2 * np.sqrt(qn) * np.sqrt(x2)
    t = mk + 2.0 * F
    t = t + np.abs(t) * 1e-6 + 1e-30
    return (a - e_up) < t[:, None]


@pytest.mark.parametrize("sample_div", [8, 1])  # 1: every tile sampled, the threshold sits right at the k-th score
@pytest.mark.parametrize("metric,dim,kind", [("euclidean", 128, "sift"), ("euclidean", 48, "gauss"), ("dot", 64, "gauss"),
                                             ("cosine", 96, "unit")])
def test_two_pass_candidates_contain_exact_topk(metric, dim, kind, sample_div):
    n, nq, k = 24_000, 64, 10
    if kind == "sift":
        X, Q = synth.sift_shaped(n, dim, 3), synth.sift_shaped(nq, dim, 4, w_seed=3)
    else:
        X = synth.latent_gaussian(n, dim, seed=dim, latent=8, normalize=(kind == "unit"))
        Q = synth.latent_gaussian(nq, dim, seed=dim + 1, w_seed=dim, latent=8, normalize=(kind == "unit"))
    keep = candidates_two_pass(X, Q, k, metric, sample_div)
    X64, Q64 = X.astype(np.float64), Q.astype(np.float64)
    if metric == "euclidean":
        d = ((Q64 ** 2).sum(1)[:, None] + (X64 ** 2).sum(1)[None, :] - 2.0 * Q64 @ X64.T)
    elif metric == "dot":
        d = -(Q64 @ X64.T)
    else:
        d = 1.0 - Q64 @ X64.T
    # everything within float32 resolution of the k-th distance counts as "could be in the exact top-k"
    kth = np.sort(d, axis=1)[:, k - 1]
    must = d <= (kth + 1e-6 * np.abs(kth) + 1e-12)[:, None]
    assert not (must & ~keep).any()
    per_query = keep.sum(1)
    assert per_query.min() >= k and per_query.max() < 4096     # CAND_CAP: nobody would fall back to the exact scan
    assert per_query.mean() < 60 * k                            # and the bound is useful, not just sound


def test_bf16_helpers():
    x = np.array([1.0, 1.00390625, 1.0078125, 3.1415927, 0.0, 65504.0], dtype=np.float32)
    r = bf16_rn(x)
    assert r[0] == 1.0 and r[2] == 1.0078125 and r[1] in (1.0, 1.0078125)   # 1 + 2^-8 is a tie: to even = 1.0
    assert r[1] == 1.0
    assert (np.abs(r - x) <= np.abs(x) * 2.0 ** -8).all()
    up = bf16_ru(x)
    assert (up >= x).all() and (up - x <= np.abs(x) * 2.0 ** -7 + 1e-30).all()
