"""Development: how fast do the k-th-best bounds tighten under different row orders? (host simulation)"""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from oracle import oracle as orc
from vaq_b200 import train

w = dict(bench.WORKLOADS["sift1m_256b_m32_k10"]); w["n"] = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000; w["nq"] = 64
pb = bench.Problem(w)
m = pb.model
om = orc.Model(m.L, m.bits, m.centroids)
port = orc.Port()
t0 = time.time(); codes = port.encode(om, pb.XP); print("encode", time.time() - t0, codes.shape)
Q = pb.Q[:64]
luts = port.create_lut(om, Q)           # [nq, lut_size]
N = codes.shape[0]; k = 10
off = om.lut_off
def dists(lo, hi, s0=0, s1=om.M):
    d = np.zeros((Q.shape[0], hi - lo), np.float32)
    for s in range(s0, s1):
        d += luts[:, off[s]:off[s + 1]][:, codes[lo:hi, s]]
    return d
SH = 125_000
D = dists(0, SH); P4 = dists(0, SH, 0, 4); P8 = dists(0, SH, 0, 8)
final_local = np.sort(D, axis=1)[:, k - 1]
if N >= 1_000_000:
    Dall = np.concatenate([np.sort(dists(i, min(N, i + SH)), axis=1)[:, :k] for i in range(0, N, SH)], axis=1)
    final_global = np.sort(Dall, axis=1)[:, k - 1]
else:
    final_global = final_local
rng = np.random.default_rng(1)
samp = rng.choice(SH, 4096, replace=False)
seed_thr = np.sort(D[:, samp], axis=1)[:, k - 1] * 1.0625
def surv(thr, P): return (P < thr[:, None]).mean()
def surv_tile(thr, P):
    a = (P < thr[:, None]).reshape(8, 8, -1).any(axis=1)       # 8 tiles of 8 queries
    return a.mean()
for name, thr in [("seed(4096 random x1.0625)", seed_thr), ("final local (125K)", final_local), ("final global (1M)", final_global)]:
    print(f"{name:32s} thr/global {np.mean(thr / final_global):.3f}  s1 per query {surv(thr, P4):.4f} per tile {surv_tile(thr, P4):.4f}   l1(8 fields) per query {surv(thr, P8):.4f} tile {surv_tile(thr, P8):.4f}  full {surv(thr, D):.5f}")
# coarse clustering on the decoded leading dims
for C_, segs in [(64, 4), (64, 8), (256, 8), (32, 4)]:
    dims = segs * m.L
    dec = np.concatenate([m.centroids[s][codes[:SH, s]] for s in range(segs)], axis=1)
    cent = train.kmeans(dec[rng.choice(SH, 20000, replace=False)], C_, iters=10)
    def assign(x): return np.argmin(((x[:, None, :] - cent[None]) ** 2).sum(-1), axis=1)
    a = np.concatenate([assign(dec[i:i + 8192]) for i in range(0, SH, 8192)])
    qa_d = ((Q[:, None, :dims] - cent[None]) ** 2).sum(-1)
    order_c = np.argsort(qa_d, axis=1)
    sizes = np.bincount(a, minlength=C_)
    for nprobe in (1, 2, 4):
        thr = np.empty(Q.shape[0], np.float32); frac = 0
        for q in range(Q.shape[0]):
            sel = np.isin(a, order_c[q, :nprobe])
            frac += sel.mean()
            dd = np.sort(D[q, sel])
            thr[q] = dd[k - 1] if dd.size >= k else np.inf
        print(f"C={C_} segs={segs} nprobe={nprobe}: rows scanned {frac / Q.shape[0]:.4f}  thr/global {np.mean(thr / final_global):.3f} thr/local {np.mean(thr / final_local):.3f}  s1 per query {surv(thr, P4):.4f}  l1 {surv(thr, P8):.4f} full {surv(thr, D):.5f}")
# trajectory in storage order: average stage-1 survival over the scan with thr(n) = k-th best of the first n rows (+ seed)
def trajectory(order, label):
    Dq = D[:, order]; P = P4[:, order]; P_8 = P8[:, order]
    tot = 0.0; tot8 = 0.0; totf = 0.0
    step = 1024
    thr = seed_thr.copy()
    best = np.full((Q.shape[0], k), np.inf, np.float32)
    for i in range(0, SH, step):
        blk = Dq[:, i:i + step]
        tot += (P[:, i:i + step] < thr[:, None]).sum(); tot8 += (P_8[:, i:i + step] < thr[:, None]).sum(); totf += (blk < thr[:, None]).sum()
        best = np.sort(np.concatenate([best, blk], axis=1), axis=1)[:, :k]
        thr = np.minimum(thr, best[:, k - 1])
    print(f"{label}: mean s1 per query over the scan {tot / D.size:.4f}  l1 {tot8 / D.size:.4f} full {totf / D.size:.5f}")
trajectory(np.arange(SH), "storage order")
