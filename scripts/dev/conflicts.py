"""Development: bank-conflict figure (wavefronts per quarter-warp and stage-1 field) of the storage order the library chose."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
import bench
from vaq_b200.index import EA, PROJECTED, VAQIndex

def quarter_conflicts(codes, order, nf=4):
    n = (order.size // 8) * 8
    r = (codes[order[:n], :nf] & 7).reshape(-1, 8, nf)
    onehot = (r[..., None] == np.arange(8)[None, None, None, :]).sum(1)
    return float(onehot.max(2).mean())

def quarter_wavefronts_exact(codes, order, nf=4):
    """wavefronts = max over bank groups of the number of DISTINCT entries mapped to it (equal entries broadcast)"""
    n = (order.size // 8) * 8
    c = codes[order[:n], :nf].astype(np.int64).reshape(-1, 8, nf)
    tot = 0.0
    for f in range(nf):
        cf = c[:, :, f]
        w = np.zeros(cf.shape[0], np.int64)
        for g in range(8):
            sel = (cf & 7) == g
            # distinct values among selected lanes
            vals = np.where(sel, cf, -1)
            vals.sort(axis=1)
            distinct = ((vals[:, 1:] != vals[:, :-1]) & (vals[:, 1:] >= 0)).sum(1) + (vals[:, 0] >= 0)
            w = np.maximum(w, distinct)
        tot += w.mean()
    return tot / nf

w = dict(bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "shard125k_256b_m32_k10"]); w["nq"] = 64
pb = bench.Problem(w)
m = pb.model
ix = VAQIndex(m.L, m.bits, m.centroids)
ix.encode_add(pb.XP)
codes = ix.get_codes()
print("arrival order:", quarter_conflicts(codes, np.arange(codes.shape[0])), quarter_wavefronts_exact(codes, np.arange(codes.shape[0])))
ix.search(pb.Q, 10, EA | PROJECTED)
order = ix.get_row_order()
print("VAQGPU_TUNE=%s storage order:" % os.environ.get("VAQGPU_TUNE", ""), quarter_conflicts(codes, order), quarter_wavefronts_exact(codes, order), ix.last_config())
