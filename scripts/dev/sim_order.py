"""Development: stage-1 survivor trajectories of a 125K-row shard under different scan orders (host simulation)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from oracle import oracle as orc
from vaq_b200 import train

NQ = 10000
NT = 40      # tiles evaluated
w = dict(bench.WORKLOADS["shard125k_256b_m32_k10"]); w["nq"] = NQ
pb = bench.Problem(w)
m = pb.model
om = orc.Model(m.L, m.bits, m.centroids)
port = orc.Port()
codes = port.encode(om, pb.XP)
Q = pb.Q[:NQ]
N = codes.shape[0]; k = 10
off = om.lut_off
rng = np.random.default_rng(1)
class Eval:
    def __init__(self, qsel):
        self.qsel = qsel
        luts = port.create_lut(om, Q[qsel])
        P4 = np.zeros((len(qsel), N), np.float32); R = np.zeros((len(qsel), N), np.float32)
        for s in range(om.M):
            v = luts[:, off[s]:off[s + 1]][:, codes[:, s]]
            if s < 4: P4 += v
            else: R += v
        self.P4 = P4; self.D = P4 + R
        self.final = np.sort(self.D, axis=1)[:, k - 1]

def run(ev, order_of_tile, tiles, label, seed_rows_of_tile=None):
    """tiles = list of (tile id, positions of its queries in ev)"""
    tot_rows = 0; tot_pairs = 0
    D, P4 = ev.D, ev.P4
    for t, qs in tiles:
        order = order_of_tile(t)
        Dq = D[qs][:, order]; Pq = P4[qs][:, order]
        if seed_rows_of_tile is not None:
            sr = seed_rows_of_tile(t)
            thr = np.sort(D[qs][:, sr], axis=1)[:, k - 1] * 1.0625
        else:
            thr = np.full(len(qs), np.inf, np.float32)
        best = np.full((len(qs), k), np.inf, np.float32)
        step = 1024
        for i in range(0, N, step):
            live = Pq[:, i:i + step] < thr[:, None]
            tot_rows += live.any(axis=0).sum(); tot_pairs += live.sum()
            best = np.sort(np.concatenate([best, Dq[:, i:i + step]], axis=1), axis=1)[:, :k]
            thr = np.minimum(thr, best[:, k - 1])
    nt = len(tiles)
    print(f"{label:40s}: stage-1 surviving rows per tile {tot_rows / nt / N:.4f}   live pairs per query {tot_pairs / (8 * nt) / N:.4f}")

sel = rng.choice(NQ // 8, NT, replace=False)
ev = Eval(np.concatenate([np.arange(8 * t, 8 * t + 8) for t in sel]))
rand_tiles = [(int(t), np.arange(8 * i, 8 * i + 8)) for i, t in enumerate(sel)]
samp = rng.choice(N, 4096, replace=False)
run(ev, lambda t: np.arange(N), rand_tiles, "arrival order, random sample seed", lambda t: samp)
live = ev.P4 < ev.final[:, None]
print(f"{'final bounds from the start':40s}: stage-1 surviving rows per tile {np.mean([live[q].any(axis=0).mean() for _, q in rand_tiles]):.4f}   live pairs per query {live.mean():.4f}")

import os
SEG_LO = int(os.environ.get("SEG_LO", "0")); SEG_N = int(os.environ.get("SEG_N", "4"))
for C_ in (64,):
    segs = SEG_N; dims = segs * m.L
    dec = np.concatenate([m.centroids[s][codes[:, s]] for s in range(SEG_LO, SEG_LO + segs)], axis=1)
    cent = train.kmeans(dec[rng.choice(N, 20000, replace=False)], C_, iters=10)
    a = np.concatenate([np.argmin(((dec[i:i + 8192, None, :] - cent[None]) ** 2).sum(-1), axis=1) for i in range(0, N, 8192)])
    qd = ((Q[:, None, SEG_LO * m.L:SEG_LO * m.L + dims] - cent[None]) ** 2).sum(-1)
    near = np.argmin(qd, axis=1)
    qorder = np.argsort(near, kind="stable")
    sel = rng.choice(NQ // 8, NT, replace=False)
    qsel = np.concatenate([qorder[8 * t:8 * t + 8] for t in sel])
    ev = Eval(qsel)
    tiles = [(int(t), np.arange(8 * i, 8 * i + 8)) for i, t in enumerate(sel)]
    first_q = {int(t): int(qorder[8 * t]) for t in sel}
    storage = np.argsort(a, kind="stable")                 # rows grouped by cluster
    start = np.concatenate([[0], np.cumsum(np.bincount(a, minlength=C_))])
    live = ev.P4 < ev.final[:, None]
    print(f"C={C_}: final bounds, grouped tiles: surviving rows per tile {np.mean([live[q].any(axis=0).mean() for _, q in tiles]):.4f}   live pairs per query {live.mean():.4f}")
    run(ev, lambda t: storage, tiles, f"C={C_} grouped tiles, storage order, random seed", lambda t: samp)
    for P in (1,):
        def order_of_tile(t, P=P):
            cl = np.argsort(qd[first_q[t]])[:P]
            if P == 1:        # rotation
                s0 = start[cl[0]]
                return np.concatenate([storage[s0:], storage[:s0]])
            probe = np.concatenate([storage[start[c]:start[c + 1]] for c in cl])
            mask = np.ones(C_, bool); mask[cl] = False
            rest = np.concatenate([storage[start[c]:start[c + 1]] for c in range(C_) if mask[c]])
            return np.concatenate([probe, rest])
        run(ev, order_of_tile, tiles, f"C={C_} P={P} probes first, seed from first 4096", lambda t: order_of_tile(t)[:4096])
    def order_all(t):
        cl = np.argsort(qd[first_q[t]])
        return np.concatenate([storage[start[c]:start[c + 1]] for c in cl])
    run(ev, order_all, tiles, f"C={C_} all clusters by proximity", lambda t: order_all(t)[:4096])
