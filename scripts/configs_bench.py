"""The other BASELINE.json shapes (C1, C3, C4 per-GPU slice, C5 per-GPU slice) on one GPU: timings + parity spot checks.
Models are trained on a host sample; the 1M+ row code matrices are generated on the device from the sample's code
distribution (vaqgpu_add_codes_synthetic) and regenerated on the host for the checks.
usage: python scripts/configs_bench.py [c1 c3 c4 c5]   -> one JSON line per config"""
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import oracle as orc
from vaq_b200 import synth, train
from vaq_b200.index import EA, PROJECTED, VAQIndex

CFG = {
    "c1": dict(desc="siftsmall-shape 10K x 128, 100 queries, VAQ 128-bit m16, k=100", n=10_000, d=128, budget=128, M=16, lo=6, hi=10,
               nq=100, k=100, train=10_000, sift=True),
    "c3": dict(desc="GIST1M-shape 1M x 960, 1K queries, VAQ 512-bit m64 (large-LUT spill path), k=10", n=1_000_000, d=960,
               budget=512, M=64, lo=4, hi=13, nq=1000, k=10, train=20_000, decay=15.0),
    "c4": dict(desc="Deep100M-shape per-GPU slice 12.5M x 96 (100M / 8), 10K queries, VAQ 128-bit m16, k=10", n=12_500_000, d=96,
               budget=128, M=16, lo=6, hi=10, nq=10_000, k=10, train=32_768, decay=4.0),
    "c5": dict(desc="1B x 128 per-GPU slice 125M rows (1B / 8), 1K queries, VAQ 256-bit m32, k=10", n=125_000_000, d=128,
               budget=256, M=32, lo=7, hi=9, nq=1000, k=10, train=32_768, decay=4.0),
}
SEED = 13517106


def run(name):
    c = CFG[name]
    t0 = time.time()
    if c.get("sift"):
        X = synth.sift_like(c["train"], c["d"], seed=SEED)
        Qraw = synth.sift_like(c["nq"], c["d"], seed=SEED + 7)
    else:
        X = synth.decaying_gaussian(c["train"], c["d"], decay=c["decay"], seed=SEED)
        Qraw = synth.decaying_gaussian(c["nq"], c["d"], decay=c["decay"], seed=SEED + 7)
    model, XP = train.train(X, c["budget"], c["M"], c["lo"], c["hi"], kmeans_iters=5, seed=SEED)
    Q = model.project(Qraw)
    om = orc.Model(model.L, model.bits, model.centroids)
    port = orc.Port()
    ix = VAQIndex(model.L, model.bits, model.centroids, eig=model.eig)
    sample_codes = port.encode(om, XP)
    if c["n"] == c["train"]:
        ix.add_codes(sample_codes)
        cdf = None
    else:
        cdf = synth.code_cdf(sample_codes, model.bits)
        ix.reserve(c["n"])
        ix.add_synthetic(c["n"], SEED, cdf)
    setup_s = time.time() - t0
    dev = torch.device("cuda", 0)
    dq = torch.from_numpy(Q).to(dev)
    lab = torch.empty((c["nq"], c["k"]), dtype=torch.int32, device=dev)
    dis = torch.empty((c["nq"], c["k"]), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ms, scan = [], []
    for i in range(2 + 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ix.search_device(dq.data_ptr(), c["nq"], c["k"], EA | PROJECTED, lab.data_ptr(), dis.data_ptr(), st)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(e0.elapsed_time(e1)); scan.append(ix.last_timings()["scan_ms"])
    cfg = ix.last_config()
    lab_h, dis_h = lab.cpu().numpy(), dis.cpu().numpy()
    # parity spot checks on a host-regenerated slice
    nchk = min(c["n"], 300_000)
    codes = sample_codes if cdf is None else synth.synth_codes(model.bits, nchk, 0, SEED, cdf)
    lut = port.create_lut(om, Q[:8])
    ok = True
    for q in range(8):
        d = port.adc_all(om, lut[q], codes)
        better = np.nonzero(d < dis_h[q, -1])[0]
        ok &= set(better.tolist()) <= set(lab_h[q].tolist())
        inside = lab_h[q] < nchk
        ok &= bool(np.array_equal(d[lab_h[q][inside]].view(np.uint32), dis_h[q][inside].view(np.uint32)))
        ok &= bool((np.diff(dis_h[q]) >= 0).all())
    if c["n"] == c["train"]:
        wl, wd = port.search_lex(om, codes, Q, c["k"])
        ok &= bool(np.array_equal(wl, lab_h) and np.array_equal(wd.view(np.uint32), dis_h.view(np.uint32)))
    T = max(1, cfg["queries_per_cta"])
    rb = ix.row_bytes
    out = {"config": name, "desc": c["desc"], "bits": model.bits.tolist(), "row_bytes": rb, "lut_entries": int(ix.lut_size),
           "search_ms": float(np.mean(ms)), "scan_ms": float(np.mean(scan)), "qps": c["nq"] / (np.mean(ms) / 1e3),
           "pairs_per_s": c["nq"] * c["n"] / (np.mean(scan) / 1e3),
           "algorithmic_GBps": (-(-c["nq"] // T) * c["n"] * rb) / (np.mean(scan) / 1e3) / 1e9, "scan_config": cfg,
           "parity_spot_check_ok": bool(ok), "setup_s": round(setup_s, 1)}
    print(json.dumps(out), flush=True)
    ix.close()


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["c1", "c3", "c4", "c5"]):
        run(n)
