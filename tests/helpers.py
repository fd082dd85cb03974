"""Shared test utilities: golden loading and the parity rules of SURVEY.md Appendix B."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

from oracle import oracle as orc  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"
RTOL = 1e-5   # BASELINE.json north_star: relative distance tolerance for ties / distances


def load_golden(name: str) -> dict:
    z = np.load(GOLDEN / f"{name}.npz", allow_pickle=True)
    return {k: z[k] for k in z.files}


def golden_model(g: dict):
    """(oracle Model, eig or None) from a golden VAQ case."""
    L = int(g["L"])
    bits = g["bits"].astype(np.int32)
    cents, off = [], 0
    for b in bits:
        K = 1 << int(b)
        cents.append(g["cent_flat"][off:off + K * L].reshape(K, L))
        off += K * L
    eig = g["eig"] if g["eig"].ndim == 2 else None
    return orc.Model(L, bits, cents), eig


def bitwise_equal(a, b) -> bool:
    a = np.ascontiguousarray(a, np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, np.float32).view(np.uint32)
    return bool(np.array_equal(a, b))


def assert_knn_equiv(lab_a, dis_a, lab_b, dis_b, rtol: float = RTOL, what: str = ""):
    """Appendix B rules 3+4: position-wise distances within rtol; ids equal except inside groups of
    (near-)equal distances, including membership ties at the k-th distance."""
    lab_a = np.asarray(lab_a); lab_b = np.asarray(lab_b)
    dis_a = np.asarray(dis_a, np.float64); dis_b = np.asarray(dis_b, np.float64)
    assert lab_a.shape == lab_b.shape, what
    tol = rtol * np.maximum(np.abs(dis_b), 1e-30)
    bad = np.abs(dis_a - dis_b) > tol
    assert not bad.any(), f"{what}: {bad.sum()} distances differ beyond rtol={rtol}: {dis_a[bad][:4]} vs {dis_b[bad][:4]}"
    nq, k = lab_a.shape
    for q in range(nq):
        if np.array_equal(lab_a[q], lab_b[q]):
            continue
        pos_b = {int(l): j for j, l in enumerate(lab_b[q])}
        for j in range(k):
            la = int(lab_a[q, j])
            if la == int(lab_b[q, j]):
                continue
            d = dis_a[q, j]
            if la in pos_b:      # present at another position: must be a tie with that position's distance
                assert abs(d - dis_b[q, pos_b[la]]) <= rtol * max(abs(d), 1e-30), f"{what}: q{q} pos{j} id {la} moved across non-tied distances"
            else:                # absent: must tie with the k-th (boundary) distance
                assert abs(d - dis_b[q, k - 1]) <= rtol * max(abs(d), 1e-30), f"{what}: q{q} pos{j} id {la} not in reference and not a boundary tie"


def assert_hamming_equiv(idx_a, dist_a, idx_b, dist_b, data=None, queries=None, what: str = ""):
    """Appendix B rule 6: distance lists identical; ids identical modulo equal-distance groups."""
    assert np.array_equal(np.asarray(dist_a), np.asarray(dist_b)), f"{what}: distance lists differ"
    idx_a = np.asarray(idx_a); idx_b = np.asarray(idx_b)
    nq, k = idx_a.shape
    for q in range(nq):
        for d in np.unique(dist_a[q]):
            sel = dist_a[q] == d
            sa, sb = set(idx_a[q][sel].tolist()), set(idx_b[q][sel].tolist())
            if sa == sb:
                continue
            # only the group at the boundary distance may differ in membership
            assert d == dist_a[q, k - 1], f"{what}: q{q} ids differ inside non-boundary distance group {d}"
        assert len(set(idx_a[q].tolist())) == k or (idx_a[q] == -1).any(), f"{what}: duplicate ids"
        if data is not None:
            valid = idx_a[q] >= 0
            x = data[idx_a[q][valid]] ^ queries[q][None, :]
            pc = np.array([[bin(int(w)).count("1") for w in row] for row in x]).sum(1)
            assert np.array_equal(pc, dist_a[q][valid]), f"{what}: returned distance does not match the row"


def hamming_lex(data: np.ndarray, queries: np.ndarray, k: int, id_base: int = 0):
    """Brute-force k smallest (distance, id) — the canonical order of the CUDA path."""
    nq = queries.shape[0]
    idx = np.full((nq, k), -1, np.int32)
    dist = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    for q in range(nq):
        x = data ^ queries[q][None, :]
        d = np.zeros(data.shape[0], np.int64)
        for w in range(x.shape[1]):
            v = x[:, w].copy()
            # popcount of uint64 via bytes table
            d += np.unpackbits(v.view(np.uint8).reshape(-1, 8), axis=1).sum(1, dtype=np.int64)
        order = np.lexsort((np.arange(d.size), d))[:k]
        idx[q, :order.size] = order + id_base
        dist[q, :order.size] = d[order]
    return idx, dist
