import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from oracle import oracle as orc
from vaq_b200 import train
w = dict(bench.WORKLOADS["shard125k_256b_m32_k10"]); w["nq"] = 8
pb = bench.Problem(w); m = pb.model
om = orc.Model(m.L, m.bits, m.centroids); port = orc.Port()
codes = port.encode(om, pb.XP); N = codes.shape[0]
rng = np.random.default_rng(1)
for segs_used, label in ((range(0, 4), "cluster on subspaces 0-3"), (range(4, 8), "cluster on subspaces 4-7"), (range(0, 8), "subspaces 0-7")):
    dec = np.concatenate([m.centroids[s][codes[:, s]] for s in segs_used], axis=1)
    cent = train.kmeans(dec[rng.choice(N, 20000, replace=False)], 64, iters=10)
    a = np.concatenate([np.argmin(((dec[i:i + 8192, None, :] - cent[None]) ** 2).sum(-1), axis=1) for i in range(0, N, 8192)])
    mx = []; nd = []
    for c in range(64):
        rows = codes[a == c]
        for f in range(4):
            h = np.bincount(rows[:, f] & 7, minlength=8) / max(1, rows.shape[0])
            mx.append(h.max()); nd.append(np.unique(rows[:, f]).size)
    print(label, "mean max residue share per (cluster, field): %.3f (uniform 0.125), distinct codes per (cluster, field): %.0f" % (np.mean(mx), np.mean(nd)))
