"""Seeded synthetic data sets in the shapes BASELINE.json names.

There is no network and the reference mount lacks siftsmall_base/learn
(.MISSING_LARGE_BLOBS), so every base set is generated; generators are
counter/seed based so CPU and GPU sides regenerate identical data.
"""
from __future__ import annotations

import numpy as np

SEED = 13517106  # echoes reference utils/Random.hpp:15


def sift_like(n: int, d: int = 128, seed: int = SEED) -> np.ndarray:
    """Non-negative integer-valued SIFT-like rows: round(clip(|N(0, sigma_j)|, 0, 255))."""
    rng = np.random.RandomState(seed % (2 ** 31))
    sigma = rng.uniform(10.0, 60.0, size=d).astype(np.float32)
    x = np.abs(rng.standard_normal((n, d)).astype(np.float32)) * sigma
    return np.round(np.clip(x, 0, 255)).astype(np.float32)


def decaying_gaussian(n: int, d: int, decay: float = 4.0, seed: int = SEED, rotate: bool = True,
                      rot_seed: int = SEED + 1, chunk: int = 1 << 18) -> np.ndarray:
    """rows = z * diag(sigma) * R, sigma_j = 1/(1 + j/decay), R a fixed random rotation."""
    rng = np.random.RandomState(seed % (2 ** 31))
    sigma = (1.0 / (1.0 + np.arange(d) / decay)).astype(np.float32)
    R = None
    if rotate:
        q, _ = np.linalg.qr(np.random.RandomState(rot_seed % (2 ** 31)).standard_normal((d, d)))
        R = q.astype(np.float32)
    out = np.empty((n, d), np.float32)
    for b in range(0, n, chunk):
        z = rng.standard_normal((min(chunk, n - b), d)).astype(np.float32) * sigma
        out[b:b + z.shape[0]] = z @ R if R is not None else z
    return out


def random_bitvectors(n: int, nbits: int, seed: int = SEED) -> np.ndarray:
    """[n, ceil(nbits/64)] uint64 words; bits above nbits in the last word are zero."""
    rng = np.random.RandomState(seed % (2 ** 31))
    w = (nbits + 63) // 64
    a = rng.randint(0, 2 ** 32, size=(n, w, 2), dtype=np.uint64)
    words = (a[..., 0] << np.uint64(32)) | a[..., 1]
    rem = nbits - (w - 1) * 64
    if rem < 64:
        words[:, -1] &= np.uint64((1 << rem) - 1)
    return np.ascontiguousarray(words)


def brute_force_knn(base: np.ndarray, queries: np.ndarray, k: int, block: int = 256) -> np.ndarray:
    """Exact squared-L2 ground truth ids [nq, k] (lowest id on ties)."""
    base = np.ascontiguousarray(base, np.float32)
    bb = (base.astype(np.float64) ** 2).sum(1)
    out = np.empty((queries.shape[0], k), np.int32)
    for b in range(0, queries.shape[0], block):
        q = queries[b:b + block].astype(np.float64)
        d = (q ** 2).sum(1)[:, None] - 2.0 * q @ base.T.astype(np.float64) + bb[None, :]
        idx = np.argsort(d, axis=1, kind="stable")[:, :k]
        out[b:b + block] = idx
    return out


def recall_at_k(labels: np.ndarray, gt: np.ndarray, k: int) -> float:
    """reference utils/Experiment.hpp:253-271 getAvgRecall: |returned ∩ gt[:K]| / K averaged."""
    labels = np.asarray(labels).reshape(-1, labels.shape[-1])[:, :k]
    hits = 0
    for i in range(labels.shape[0]):
        hits += len(set(labels[i].tolist()) & set(gt[i, :k].tolist()))
    return hits / (labels.shape[0] * k)


# ---- counter-based generators restated from the device kernels (vaq_b200/csrc/pack.cu
# synth_codes_kernel, hamming_scan.cu ham_synth_kernel) so any row slice of a 100M / 1B-row
# on-device index can be regenerated on the host for parity checks.
_G1 = np.uint64(0x9E3779B97F4A7C15)
_G2 = np.uint64(0xD1B54A32D192ED03)


def _mix64(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return x


def synth_codes(bits, n: int, row0: int, seed: int, cdf=None) -> np.ndarray:
    """[n, M] uint16 codes of global rows [row0, row0+n) exactly as vaqgpu_add_codes_synthetic makes them."""
    bits = np.asarray(bits, np.int64)
    M = bits.size
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
    subs = np.arange(M, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        h = _mix64(np.uint64(seed) ^ (rows * _G1 + subs * _G2))
    top = (h >> np.uint64(40)).astype(np.uint32)
    K = (1 << bits)
    if cdf is None:
        return (top & (K - 1).astype(np.uint32)[None, :]).astype(np.uint16)
    cdf = np.asarray(cdf, np.float32)
    off = np.concatenate([[0], np.cumsum(K)])
    u = top.astype(np.float32) * np.float32(1.0 / 16777216.0)
    out = np.empty((n, M), np.uint16)
    for s in range(M):
        t = cdf[off[s]:off[s + 1]]
        out[:, s] = np.minimum(np.searchsorted(t, u[:, s], side="right"), K[s] - 1).astype(np.uint16)
    return out


def synth_bitvectors(n: int, row0: int, nbits: int, seed: int) -> np.ndarray:
    """[n, ceil(nbits/64)] uint64 words of global rows [row0, row0+n) as hamgpu_add_synthetic makes them."""
    w64 = (nbits + 63) // 64
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
    ws = np.arange(w64, dtype=np.uint64)[None, :]
    with np.errstate(over="ignore"):
        x = _mix64(np.uint64(seed) ^ (rows * _G1 + ws * _G2))
    rem = nbits - (w64 - 1) * 64
    if rem < 64:
        x[:, -1] &= np.uint64((1 << rem) - 1)
    return np.ascontiguousarray(x)


def code_cdf(codes: np.ndarray, bits) -> np.ndarray:
    """Empirical per-subspace cumulative code distribution (concatenated, last entry of each table = 1)."""
    bits = np.asarray(bits, np.int64)
    out = []
    for s in range(bits.size):
        K = 1 << int(bits[s])
        cnt = np.bincount(codes[:, s].astype(np.int64), minlength=K).astype(np.float64)
        c = np.cumsum(cnt) / cnt.sum()
        c[-1] = 1.0
        out.append(c.astype(np.float32))
    return np.concatenate(out)
