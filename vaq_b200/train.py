"""Host-side VAQ training (stays on the CPU, as the north star specifies).

Restates reference ``VAQ::train`` (bitvecengine/VAQ.cpp:11-661) with numpy:

* uncentred second-moment matrix ``X^T X`` on at most ``1000*D`` sampled rows
  (VAQ.cpp:16-59), eigen-decomposition, eigenpairs sorted by eigenvalue descending
  (:84-100);
* ``L = ceil(D / M)`` dims per subspace (:102-106); partial variance balancing by
  swapping eigen-columns ``i`` and ``i*L + L-1`` while the per-subspace variance stays
  descending (:262-280);
* projection of the training rows (:294);
* variance-driven bit allocation: the ILP of VAQ.cpp:339-452
  (max sum var_s*x_s, sum x = budget, min<=x_s<=max, x_s - x_{s+1} <= nextPow2(var_s/var_{s+1}))
  solved exactly by dynamic programming (GLPK is a third-party solver the reference
  links but does not vendor; any optimal solution is acceptable to the reference);
* one Lloyd k-means codebook per subspace, 25 iterations, deterministic subset
  seeding (the reference calls ``arma::kmeans(..., static_subset, 25)``, :627 —
  Armadillo is likewise un-vendored).

Training parity with the reference is *unpinned* (no reference test covers it and
GLPK/Armadillo are absent); search parity does not depend on it because the oracle
and the GPU path consume the same trained model.
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

import numpy as np

SEED = 13517106  # reference utils/Random.hpp:15


@dataclass
class VAQModel:
    """Trained state == the reference's public members VAQ.hpp:57-73."""
    L: int                                   # mSubsLen
    bits: np.ndarray                         # mBitsAlloc [M]
    centroids: list                          # mCentroidsPerSubs, [2^bits[s], L] float32 each
    eig: np.ndarray | None = None            # real(mEigenVectors) [D, D] (None => identity)
    var_per_subs: np.ndarray | None = None   # normalised variance per subspace
    extra: dict = field(default_factory=dict)

    @property
    def M(self) -> int:
        return int(len(self.bits))

    @property
    def D(self) -> int:
        return self.M * self.L

    def project(self, X: np.ndarray) -> np.ndarray:
        """``(X * mEigenVectors).real()`` — reference VAQ.hpp:198-201 (host GEMM)."""
        X = np.ascontiguousarray(X, np.float32)
        if X.shape[1] < self.D:                       # demo_vaq.cpp:66-83 zero-pads to M*L
            X = np.pad(X, ((0, 0), (0, self.D - X.shape[1])))
        return X if self.eig is None else np.ascontiguousarray(X @ self.eig, np.float32)


def parse_method_string(s: str) -> dict:
    """``VAQ<budget>m<M>min<a>max<b>var<v>,<MODE>[_TI<c>[m<seg>]]`` — reference
    VAQ::parseMethodString, VAQ.cpp:1189-1267."""
    out = dict(budget=None, M=None, min_bits=None, max_bits=None, var=1.0, methods=set(),
               ti_clusters=None, ti_segments=-1)
    for tok in s.split(","):
        m = re.match(r"VAQ(\d+)m(\d+)min(\d+)max(\d+)var([0-9.]+)", tok)
        if m:
            out.update(budget=int(m[1]), M=int(m[2]), min_bits=int(m[3]), max_bits=int(m[4]), var=float(m[5]))
            continue
        for part in tok.split("_"):
            if "SORT" in part:
                out["methods"].add("SORT")
            elif "HEAP" in part:
                out["methods"].add("HEAP")
            elif "EA" in part:
                out["methods"].add("EA")
            elif "TI" in part:
                out["methods"].add("TI")
                m2 = re.match(r"TI(\d+)m(\d+)", part) or re.match(r"TI(\d+)", part)
                if m2:
                    out["ti_clusters"] = int(m2[1])
                    if m2.lastindex and m2.lastindex >= 2:
                        out["ti_segments"] = int(m2[2])
            elif "FAST" in part:
                raise ValueError("FAST* scan modes are out of scope (lossy uint8 LUT, SURVEY §2 #12)")
    return out


def next_pow2(x: float) -> int:
    """reference utils/Math.hpp:182-187"""
    if x == 0 or not math.isfinite(x):
        return 0
    return int(2 ** math.floor(math.log2(abs(x))))


def allocate_bits(var_per_subs: np.ndarray, budget: int, min_bits: int, max_bits: int) -> np.ndarray:
    """Exact DP for the ILP of VAQ.cpp:339-452 (all subspaces inside the variance cut, var=1)."""
    v = np.asarray(var_per_subs, np.float64)
    M = v.size
    if not (M * min_bits <= budget <= M * max_bits):
        raise ValueError(f"budget {budget} infeasible for M={M}, min={min_bits}, max={max_bits}")
    kdiff = [max(next_pow2(v[i] / v[i + 1]), 0) for i in range(M - 1)]
    NEG = -1e300
    nx = max_bits - min_bits + 1
    # best[s][xi][r] = max objective of subspaces s..M-1 given x_s = min+xi and r bits for s..M-1
    best = np.full((M, nx, budget + 1), NEG)
    choice = np.zeros((M, nx, budget + 1), np.int16)
    for xi in range(nx):
        x = min_bits + xi
        if x <= budget:
            best[M - 1, xi, x] = v[M - 1] * x
    for s in range(M - 2, -1, -1):
        for xi in range(nx):
            x = min_bits + xi
            cand_best = np.full(budget + 1, NEG)
            cand_arg = np.zeros(budget + 1, np.int16)
            for yi in range(nx):
                y = min_bits + yi
                if x - y > kdiff[s]:
                    continue
                val = np.full(budget + 1, NEG)
                val[x:] = best[s + 1, yi, :budget + 1 - x] + v[s] * x
                upd = val > cand_best
                cand_best[upd] = val[upd]
                cand_arg[upd] = yi
            best[s, xi] = cand_best
            choice[s, xi] = cand_arg
    xi = int(np.argmax(best[0, :, budget]))
    if best[0, xi, budget] <= NEG / 2:
        raise ValueError("bit allocation infeasible under the neighbour-difference constraints")
    bits = np.zeros(M, np.int32)
    r = budget
    for s in range(M):
        bits[s] = min_bits + xi
        nxt = int(choice[s, xi, r]) if s < M - 1 else 0
        r -= int(bits[s])
        xi = nxt
    assert bits.sum() == budget
    return bits


def kmeans(X: np.ndarray, K: int, iters: int = 25) -> np.ndarray:
    """Lloyd, deterministic subset seeding; empty clusters keep their previous centre."""
    X = np.ascontiguousarray(X, np.float32)
    n = X.shape[0]
    if n < K:                                         # degenerate: more centroids than samples
        reps = int(math.ceil(K / n))
        X = np.tile(X, (reps, 1))[:max(K, n)] + np.float32(1e-6) * np.arange(max(K, n), dtype=np.float32)[:, None]
        n = X.shape[0]
    C = X[(np.arange(K, dtype=np.int64) * n) // K].copy()
    blk = max(1, (1 << 20) // max(K, 1))              # 4 MB distance blocks stay in cache
    for _ in range(iters):
        cc = (C * C).sum(1)
        m2ct = np.ascontiguousarray((np.float32(-2.0) * C).T)
        assign = np.empty(n, np.int64)
        for b in range(0, n, blk):
            d = X[b:b + blk] @ m2ct                   # |x|^2 is constant per row: argmin of |c|^2 - 2 x.c
            d += cc[None, :]
            assign[b:b + blk] = d.argmin(1)
        cnt = np.bincount(assign, minlength=K)
        newC = np.zeros_like(C, dtype=np.float64)
        for j in range(X.shape[1]):
            newC[:, j] = np.bincount(assign, weights=X[:, j], minlength=K)
        nz = cnt > 0
        newC[nz] /= cnt[nz, None]
        newC[~nz] = C[~nz]
        newC = newC.astype(np.float32)
        if np.array_equal(newC, C):
            break
        C = newC
    return np.ascontiguousarray(C, np.float32)


def train(X: np.ndarray, budget: int, M: int, min_bits: int, max_bits: int, *, pca: bool = True,
          kmeans_iters: int = 25, sample_per_centroid: int = 256, max_sample: int = 1 << 18,
          seed: int = SEED):
    """Returns ``(model, X_projected)``.  The reference projects its argument in place
    (VAQ.cpp:294); here the projected rows are returned for ``encode``."""
    X = np.ascontiguousarray(X, np.float32)
    n, D0 = X.shape
    L = -(-D0 // M)
    D = M * L
    if D != D0:
        X = np.pad(X, ((0, 0), (0, D - D0)))
    rng = np.random.RandomState(seed % (2 ** 31))
    if pca:
        ns = min(n, 1000 * D)
        S = X if ns == n else X[rng.permutation(n)[:ns]]
        cov = (S.astype(np.float64).T @ S.astype(np.float64))
        w, V = np.linalg.eigh(cov)
        order = np.argsort(-w, kind="stable")
        w, V = w[order], V[:, order]
        # partial balancing, VAQ.cpp:262-280
        def subs_sorted(wv):
            e = wv.reshape(M, L).sum(1)
            return bool(np.all(e[:-1] > e[1:]))
        for i in range(1, min(L, M)):
            j = i * L + (L - 1)
            w[[i, j]] = w[[j, i]]
            if not subs_sorted(w):
                w[[i, j]] = w[[j, i]]
                break
            V[:, [i, j]] = V[:, [j, i]]
        eig = np.ascontiguousarray(V, np.float32)
        XP = np.ascontiguousarray(X @ eig, np.float32)
        var_dim = np.maximum(w / w.sum(), 1e-12)
    else:
        eig = None
        XP = X
        var_dim = (XP.astype(np.float64) ** 2).mean(0)
        var_dim = np.maximum(var_dim / var_dim.sum(), 1e-12)
    var_subs = var_dim.reshape(M, L).sum(1)
    bits = allocate_bits(var_subs, budget, min_bits, max_bits)
    cents = []
    for s in range(M):
        K = 1 << int(bits[s])
        ns = min(n, max(K * sample_per_centroid, sample_per_centroid << (budget // M)), max_sample)
        idx = rng.permutation(n)[:ns] if ns < n else np.arange(n)
        cents.append(kmeans(XP[idx, s * L:(s + 1) * L], K, kmeans_iters))
    model = VAQModel(L=L, bits=bits, centroids=cents, eig=eig, var_per_subs=var_subs)
    return model, XP

