"""One workload per invocation, a handful of launches — the target of the ncu captures under profiles/.
usage: python scripts/prof_run.py {sift1m | hbm1 | hbm8 | hbm64 | ham1 | ham2 | ham8 | ham64} [reps]
ncu is wrapped around it, e.g.
  ncu --set full --clock-control none --import-source on -k regex:adc_filter16 -s 2 -c 1 -o gpurun_out/x python scripts/prof_run.py sift1m
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402  (workload definitions and the seeded problem builder)
from vaq_b200 import synth  # noqa: E402
from vaq_b200.index import EA, PROJECTED, HammingIndex, VAQIndex  # noqa: E402

case = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
k = 10
if case.startswith("ham"):
    nq = int(case[3:])
    hx = HammingIndex(256)
    hx.add_synthetic(64_000_000, bench.SEED)
    q = torch.from_numpy(synth.synth_bitvectors(nq, 10 ** 10, 256, bench.SEED).view(np.int64)).to(dev)
    idx = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dist = torch.empty((nq, k), dtype=torch.int32, device=dev)
    for _ in range(reps):
        hx.query_device(q.data_ptr(), nq, k, idx.data_ptr(), dist.data_ptr(), st)
        torch.cuda.synchronize()
    print(case, hx.last_timings(), hx.last_config())
else:
    w = dict(bench.WORKLOADS["sift1m_256b_m32_k10"])
    if case != "sift1m":
        w["n"] = 200_000           # only the model and the code distribution are needed
    pb = bench.Problem(w)
    ix = VAQIndex(pb.model.L, pb.model.bits, pb.model.centroids)
    ix.encode_add(pb.XP)
    nq = pb.nq
    if case != "sift1m":
        nq = int(case[3:])
        cdf = synth.code_cdf(ix.get_codes(), pb.model.bits)
        ix.close()
        ix = VAQIndex(pb.model.L, pb.model.bits, pb.model.centroids)
        ix.reserve(64_000_000)
        ix.add_synthetic(64_000_000, bench.SEED, cdf)
    dq = torch.from_numpy(pb.Q[:nq]).to(dev)
    lab = torch.empty((nq, k), dtype=torch.int32, device=dev)
    dis = torch.empty((nq, k), dtype=torch.float32, device=dev)
    for _ in range(reps):
        ix.search_device(dq.data_ptr(), nq, k, EA | PROJECTED, lab.data_ptr(), dis.data_ptr(), st)
        torch.cuda.synchronize()
    print(case, ix.last_timings(), ix.last_config())
