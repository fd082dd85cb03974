"""Hamming-scan micro-benchmark (for ncu): python scripts/ham_bench.py [rows] [nq] [k]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from vaq_b200 import synth
from vaq_b200.index import HammingIndex

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 64_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 64
k = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda", 0)
hx = HammingIndex(256)
hx.add_synthetic(rows, 1)
q = torch.from_numpy(synth.synth_bitvectors(nq, 10 ** 10, 256, 1).view(np.int64)).to(dev)
idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
dist = torch.empty((nq, k), dtype=torch.int32, device=dev)
st = torch.cuda.current_stream().cuda_stream
ms = []
for i in range(6):
    hx.query_device(q.data_ptr(), nq, k, idx.data_ptr(), dist.data_ptr(), st)
    torch.cuda.synchronize()
    ms.append(hx.last_timings()["scan_ms"])
print("scan_ms", ms, hx.last_config(), "pairs/s %.3g" % (rows * nq / (min(ms) / 1e3)))
