#!/usr/bin/env python
"""Headline benchmark: VAQ query-time search (LUT build -> ADC scan -> top-k) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU search

Default workload (BASELINE.json configs[1]): SIFT1M-shape synthetic, 1M x 128 base, 10K queries, VAQ 256-bit budget
over 32 subspaces, k = 10.  A step = one search of the whole query batch against the index.  `--workload` selects
the other BASELINE shapes (siftsmall C1, GIST1M C3 spill path, Deep100M C4, 1B C5).

N > 1 (torchrun, one rank per GPU): the packed code matrix is ROW-SHARDED over the N GPUs (rank r holds rows
[r*ceil(n/N), (r+1)*ceil(n/N))), every rank scans its rows for all queries while the shards exchange their running
k-th-best bounds through NVLink peer memory, then one NCCL all-gather of the shard-local top-k key lists + a device
merge.  Strong scaling: total rows and queries fixed.  (`--row-shards R` < N is an experiment knob: N/R replica
groups of R row shards that split the query batch.)

One JSON line on stdout (rank 0); everything else goes to stderr.
"""
from __future__ import annotations

import os

if "LOCAL_RANK" in os.environ and int(os.environ.get("WORLD_SIZE", "1")) > 1:
    # torchrun exports OMP_NUM_THREADS=1; the host-side problem set-up (numpy) may use this rank's share of the cores
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ["WORLD_SIZE"])))

import argparse
import json
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

# source "vectors": base vectors generated on the host and encoded on the device (VAQ::encode semantics);
# source "codes": model trained on a host sample, rows generated directly on the device from the sample's per-subspace
#                 code distribution (the 100M / 1B-row shapes no host can hold), regenerated on the host
#                 (vaq_b200/synth.py) for any row range the checks need.
WORKLOADS = {
    "sift1m_256b_m32_k10": dict(n=1_000_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=10_000, k=10, decay=4.0,
                                source="vectors", desc="BASELINE configs[1] (C2): SIFT1M-shape 1M x 128, 10K queries, VAQ 256-bit m32, k=10"),
    "siftsmall_128b_m16_k100": dict(n=10_000, d=128, budget=128, M=16, min_bits=6, max_bits=10, nq=100, k=100, decay=4.0,
                                    source="vectors", sift=True, train_rows=10_000,
                                    desc="BASELINE configs[0] (C1): siftsmall-shape 10K x 128, the reference's 100 shipped queries, VAQ 128-bit, k=100"),
    "sift1m_256b_m32_min2max13_k10": dict(n=1_000_000, d=128, budget=256, M=32, min_bits=2, max_bits=13, nq=10_000, k=10, decay=4.0,
                                          source="vectors",
                                          desc="the reference's own SIFT1M setting (ExperimentsParameters.txt:55: 256 bit, 32 segments, min 2 / max 13 bits)"),
    "sift1m_256b_m32_ti1000_k10": dict(n=1_000_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=10_000, k=10, decay=4.0,
                                       source="vectors", ti_clusters=1000, ti_segments=16, visit=0.25,
                                       desc="C2 model with the reference's SIFT1M TI setting (ExperimentsParameters.txt:55): 1000 TI clusters "
                                            "over 16 segments, visit 25 % of the clusters; clustering on the device (vaqgpu_cluster_ti)"),
    "gist1m_512b_m64_k10": dict(n=1_000_000, d=960, budget=512, M=64, min_bits=4, max_bits=13, nq=1_000, k=10, decay=15.0,
                                source="codes", train_rows=20_000,
                                desc="BASELINE configs[2] (C3): GIST1M-shape 1M x 960, 1K queries, VAQ 512-bit m64 min4/max13 (large-LUT spill path)"),
    "deep100m_128b_m16_k10": dict(n=100_000_000, d=96, budget=128, M=16, min_bits=6, max_bits=10, nq=10_000, k=10, decay=4.0,
                                  source="codes", desc="BASELINE configs[3] (C4): Deep100M-shape 100M x 96, 10K queries, VAQ 128-bit m16"),
    "synth1b_256b_m32_k10": dict(n=1_000_000_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=1_000, k=10, decay=4.0,
                                 source="codes", ham_rows=1_000_000_000,
                                 desc="BASELINE configs[4] (C5): 1B x 128, VAQ 256-bit m32, 1K queries per step + 1B x 256-bit Hamming scan"),
    # development shapes
    "small_256b_m32_k10": dict(n=100_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=1_000, k=10, decay=4.0, source="vectors",
                               desc="development"),
    "shard125k_256b_m32_k10": dict(n=125_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=10_000, k=10, decay=4.0, source="vectors",
                                   desc="development: one of eight row shards of the SIFT1M shape"),
}
TRAIN_ROWS = 32768
SEED = 13517106
FP32_FMA_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12     # 148 SMs x 128 FP32 lanes x 2 flop x 1.965 GHz


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly one JSON line: everything else any library prints at the C level (e.g. NCCL's version
# banner) is sent to stderr by pointing fd 1 at fd 2 for the whole run and keeping the real stdout aside.
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


class Problem:
    """Seeded synthetic base/queries + host-trained model (the reference keeps training on the host)."""

    def __init__(self, w: dict, need_vectors: bool = True):
        from vaq_b200 import io as vio
        from vaq_b200 import synth, train
        t0 = time.time()
        self.w = w
        self.n, self.nq, self.k = w["n"], w["nq"], w["k"]
        train_rows = w.get("train_rows", TRAIN_ROWS)
        gen = (lambda n, seed: synth.sift_like(n, w["d"], seed=seed)) if w.get("sift") else \
              (lambda n, seed: synth.decaying_gaussian(n, w["d"], decay=w["decay"], seed=seed))
        self.X = self.XP = None
        if w["source"] == "vectors":
            self.X = gen(self.n, SEED)
            Xtrain = self.X[:train_rows]
        else:
            Xtrain = gen(train_rows, SEED)
        qfix = ROOT / "tests" / "golden" / "siftsmall_query.fvecs"
        if w.get("sift") and qfix.exists():
            self.Qraw = vio.read_fvecs(qfix)[: self.nq]          # the reference's shipped siftsmall queries (fixture)
        else:
            self.Qraw = gen(self.nq, SEED + 7)
        self.model, XPtrain = train.train(Xtrain, w["budget"], w["M"], w["min_bits"], w["max_bits"], kmeans_iters=8, seed=SEED)
        self.Q = self.model.project(self.Qraw)
        self.cdf = None
        if w["source"] == "vectors":
            if need_vectors:
                self.XP = self.model.project(self.X)
        else:
            self.XPtrain = XPtrain
        log(f"[bench] problem built in {time.time() - t0:.1f}s bits={self.model.bits.tolist()}")

    def make_cdf(self, sample_codes):
        from vaq_b200 import synth
        self.cdf = synth.code_cdf(sample_codes, self.model.bits)

    def host_codes(self, row0: int, n: int) -> np.ndarray:
        """codes of global rows [row0, row0+n) of a "codes" workload, regenerated on the host"""
        from vaq_b200 import synth
        return synth.synth_codes(self.model.bits, n, row0, SEED, self.cdf)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            p = [x.strip() for x in r.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); smax.append(float(p[1])); pw.append(float(p[2]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak_hbm() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
        except Exception:
            pass
    return 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"


def ncu_traffic(kernel: str, workload: str, n_gpus: int):
    """DRAM bytes per launch of `kernel` on `workload` from the committed ncu --set full captures (profiles/ncu_traffic.json)."""
    try:
        tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        for ent in tj.get(kernel, []) if isinstance(tj.get(kernel), list) else [tj.get(kernel, {})]:
            if ent.get("workload") == workload and ent.get("n_gpus", 1) == n_gpus:
                return ent.get("dram_bytes_read", 0) + ent.get("dram_bytes_write", 0)
    except Exception:
        pass
    return None


def use_host_threads(n: int):
    try:        # torchrun exports OMP_NUM_THREADS=1; the CPU checkers may use the idle host cores
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(max(1, n))
    except OSError:
        pass


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU search (compiled unmodified reference when available)
# ---------------------------------------------------------------------------------------------

def cpu_reference_search(model, codes, Q, k, seconds_target: float, threads: int, ti: dict | None = None):
    """Times reference VAQ::search (EA mode, the reference's fastest exact mode; TI + EA over the given clusters when
    `ti` is set) on a bounded sample of the query batch against `codes`.  Returns (qps, n_queries, kind, labels,
    dists, seconds)."""
    from oracle import oracle as orc
    om = orc.Model(model.L, model.bits, model.centroids)
    if ti is not None and not orc.Ref.available():
        raise RuntimeError("the TI baseline needs the compiled reference (oracle/_ref)")
    if orc.Ref.available():
        kind = "reference"
        rv = orc.Ref().vaq(om, orc.NN_EA | (orc.NN_TI if ti is not None else 0))
        rv.set_codes(codes)
        if ti is not None:
            rv.set_ti(ti["clusters"], ti["start"], ti["sizes"], ti["members"], ti["code_to_cc"])
            rv.set_visit(ti["visit"])

        def run(q):
            return rv.search(q, k, nthreads=threads)
    else:
        kind = "port"
        port = orc.Port()

        def run(q):
            return port.search(om, codes, q, k, "EA", nthreads=threads)
    probe = min(Q.shape[0], max(threads, 16))
    t0 = time.perf_counter()
    run(Q[:probe])
    t_probe = time.perf_counter() - t0
    n = int(min(Q.shape[0], max(probe, seconds_target / max(t_probe / probe, 1e-9))))
    n = max(threads, (n // threads) * threads)
    n = min(n, Q.shape[0])
    t0 = time.perf_counter()
    lab, dis = run(Q[:n])
    dt = time.perf_counter() - t0
    return n / dt, n, kind, lab, dis, dt


def reference_codes(pb: Problem, max_rows: int):
    """(codes, rows_used, how) for the CPU arm: the whole matrix encoded by the reference itself for "vectors"
    workloads, a host-regenerated row slice for "codes" workloads."""
    from oracle import oracle as orc
    om = orc.Model(pb.model.L, pb.model.bits, pb.model.centroids)
    if pb.w["source"] == "vectors":
        t0 = time.time()
        if orc.Ref.available():      # VAQ::encode of the compiled reference (OpenMP over rows, VAQ.cpp:733)
            rv = orc.Ref().vaq(om, orc.NN_EA)
            codes = rv.encode(pb.XP)
            rv.close()
        else:
            codes = orc.Port().encode(om, pb.XP)
        log(f"[bench] host encode {time.time() - t0:.1f}s")
        return codes, pb.n, "all rows"
    if pb.cdf is None:
        pb.make_cdf(orc.Port().encode(om, pb.XPtrain))
    rows = min(pb.n, max_rows)
    return pb.host_codes(0, rows), rows, f"rows [0,{rows}) of {pb.n} (QPS scaled linearly by {rows}/{pb.n})"


def run_reference_arm(args, w, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pb = Problem(w)
    threads = os.cpu_count() or 1
    use_host_threads(threads)
    codes, rows_used, how = reference_codes(pb, args.cpu_rows)
    scale = rows_used / pb.n
    vals = []
    n_used = 0
    kind = "reference"
    per_step_target = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for it in range(args.warmup + args.steps):
        qps, n_used, kind, _, _, dt = cpu_reference_search(pb.model, codes, pb.Q, pb.k, per_step_target, threads)
        if it >= args.warmup:
            vals.append((qps * scale, dt))
    qps = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": "queries/sec at recall@10", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(name, w, "EA (VAQ::searchEarlyAbandon)"),
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": f"{n_used} of {w['nq']} queries per step, {how}, query-sliced over {threads} threads"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def workload_config(name, w, mode):
    return {"workload": name, "rows": w["n"], "dims": w["d"], "queries": w["nq"], "k": w["k"], "bits": w["budget"],
            "subspaces": w["M"], "mode": mode}


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------

def run_gpu_arm(args, w, name):
    import torch
    import torch.distributed as dist
    from vaq_b200 import synth
    from vaq_b200.index import EA, PROJECTED, HammingIndex, VAQIndex
    from vaq_b200.sharded import ShardedHamming, ShardedVAQ

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        log(f"[bench] WORLD_SIZE={world} overrides --gpus {args.gpus}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    pb = Problem(w)
    model, Q = pb.model, pb.Q
    n, nq, k = pb.n, pb.nq, pb.k
    flags = EA | PROJECTED | (0x1000 if args.scan_v1 else 0)

    # ---- index: this rank's row block.  Default R = N: the pure row-sharded layout BASELINE.json names.
    R = args.row_shards or world
    sh = ShardedVAQ(model.L, model.bits, model.centroids, model.eig, n, rank, world, local_rank, row_shards=R)
    t0 = time.time()
    if w["source"] == "vectors":
        sh.index.encode_add(pb.XP[sh.lo:sh.hi])          # device encode (bit-exact vs VAQ::encode, checked below)
    else:
        # the training sample, encoded on the device, only shapes the synthetic code distribution
        tmp = VAQIndex(model.L, model.bits, model.centroids, device=local_rank)
        tmp.encode_add(pb.XPtrain)
        pb.make_cdf(tmp.get_codes())
        tmp.close()
        sh.add_synthetic(SEED, pb.cdf)
    torch.cuda.synchronize()
    log(f"[bench] rank {rank}: rows [{sh.lo},{sh.hi}) resident in {time.time() - t0:.1f}s; row_bytes={sh.index.row_bytes}")
    ix = sh.index
    if w.get("ti_clusters"):
        if world > 1:
            raise SystemExit("the TI workload runs on one GPU in this bench (sharded TI: tests/test_gpu_sharded.py)")
        from vaq_b200.index import SQRT, TI
        t0 = time.time()
        ix.cluster_ti(w["ti_clusters"], w["ti_segments"], 10)
        ix.set_visit(w["visit"])
        flags = TI | EA | SQRT | PROJECTED
        log(f"[bench] device clusterTI: {w['ti_clusters']} clusters over {w['ti_segments']} segments in {time.time() - t0:.2f}s")
    exchange = False
    if world > 1 and not args.no_bound_exchange:
        exchange = sh.enable_bound_exchange(nq)

    d_q = torch.from_numpy(Q).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream()

    def step_device():
        return sh.search(d_q, k, flags)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    # started before warm-up: NVML start-up stays out of the timed steps (VAQ_BENCH_NO_SAMPLER: development, to rule the sampler out)
    sampler = ClockSampler(local_rank) if rank == 0 and not os.environ.get("VAQ_BENCH_NO_SAMPLER") else None
    for _ in range(args.warmup):
        labels, dists = step_device()      # results held like in the timed steps: the caching allocator reaches its steady state here
    barrier()
    if sampler:
        sampler.rows.clear()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    scan_ms, lut_ms, merge_ms = [], [], []
    launches = 0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)            # L2 flush between timed steps (outside the per-step event bracket)
        barrier()
        ev[i][0].record(st)
        labels, dists = step_device()
        ev[i][1].record(st)
        torch.cuda.synchronize()
        t = ix.last_timings()
        scan_ms.append(t["scan_ms"]); lut_ms.append(t["lut_ms"]); merge_ms.append(t["merge_ms"])
        launches += ix.last_config()["launches"] + (1 if world > 1 else 0) + sh.qgroups   # + all-gather + final merge(s)
    barrier()
    clocks = sampler.stop() if sampler else None
    step_ms = np.array([a.elapsed_time(b) for a, b in ev], dtype=np.float64)
    total_ms = torch.tensor([step_ms.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = nq * args.steps / (total_ms / 1e3)
    scan_all = torch.tensor([float(np.mean(scan_ms)), float(np.mean(lut_ms)), float(np.mean(merge_ms))], dtype=torch.float64, device=dev)
    scan_max = scan_all.clone()
    if world > 1:
        dist.all_reduce(scan_max, op=dist.ReduceOp.MAX)
    scan_max = [float(x) for x in scan_max.cpu()]

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region
    q_pin = torch.from_numpy(Q).pin_memory()
    lab_pin = torch.empty((nq, k), dtype=torch.int32).pin_memory()
    dis_pin = torch.empty((nq, k), dtype=torch.float32).pin_memory()

    def step_e2e():
        if world == 1:
            ix.search_into(q_pin.numpy(), k, flags, lab_pin.numpy(), dis_pin.numpy())     # the C-ABI host call
        else:
            dq = q_pin.to(dev, non_blocking=True)
            l, d = sh.search(dq, k, flags)
            lab_pin.copy_(l, non_blocking=True)
            dis_pin.copy_(d, non_blocking=True)
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_qps = nq * args.steps / float(e2e_s.item())

    peak, peak_src = measured_peak_hbm()

    # ---- planted needles (device-generated workloads; collective): queries placed on the code tuple of known rows,
    # one or more per shard, must return that row first with the oracle's distance (checked on rank 0 below)
    needles = needle_q = needle_out = None
    if w["source"] == "codes":
        needles = [int((r + 0.37) * n / max(world, 8)) for r in range(max(world, 8))]
        needle_q = np.stack([np.concatenate([model.centroids[s][c] for s, c in enumerate(pb.host_codes(p, 1)[0])])
                             for p in needles]).astype(np.float32)
        nl, nd = sh.search(torch.from_numpy(needle_q).to(dev), k, flags)
        needle_out = (nl.cpu().numpy(), nd.cpu().numpy())

    # ---- Hamming scan, row-sharded over the same ranks (collective: every rank takes part)
    hamming_sharded = None
    if not args.no_hamming:
        try:
            hn = int(w.get("ham_rows", args.ham_rows))
            hs = ShardedHamming(256, hn, rank, world, local_rank)
            hs.add_synthetic(SEED)
            hq_n = args.ham_queries
            # planted needles: query j is row p_j of the matrix with 3 bits flipped -> must come back first at distance 3
            planted = (np.arange(min(8, hq_n), dtype=np.int64) * (hn // 8) + 12345) % hn
            hq_host = synth.synth_bitvectors(hq_n, 10 ** 10, 256, SEED)
            for j, p in enumerate(planted):
                row = synth.synth_bitvectors(1, int(p), 256, SEED)[0]
                row[0] ^= np.uint64(0b111)
                hq_host[j] = row
            hqv = torch.from_numpy(hq_host.view(np.int64)).to(dev)
            hms = []
            for i in range(3 + 5):
                flush.fill_(i)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                hidx, hdist = hs.query(hqv, k)
                e1.record(st)
                torch.cuda.synchronize()
                if i >= 3:
                    hms.append(e0.elapsed_time(e1))
            tm = torch.tensor([float(np.sum(hms))], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            hscan = hs.index.last_timings()["scan_ms"]
            hcfg = hs.index.last_config()
            qt = max(1, hcfg["queries_per_cta"])
            n_loc = hs.hi - hs.lo
            hi_h, hd_h = hidx.cpu().numpy(), hdist.cpu().numpy()
            needles_ok = bool(all(hi_h[j, 0] == planted[j] and hd_h[j, 0] == 3 for j in range(len(planted))))
            hbytes = -(-hq_n // qt) * n_loc * 32
            hamming_sharded = {"kernel": "ham_scan_kernel", "rows_total": hn, "rows_per_gpu": n_loc, "bits": 256, "queries": hq_n,
                               "k": k, "sharding": f"rows/{world}", "qps": hq_n * 5 / (float(tm.item()) / 1e3),
                               "ms_per_batch": float(tm.item()) / 5, "scan_ms_rank0": hscan, "queries_per_pass": qt,
                               "achieved": hbytes / (hscan / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                               "frac": hbytes / (hscan / 1e3) / 1e9 / peak, "planted_needles_ok": needles_ok}
            hs.index.close()
        except Exception as e:      # never hide the headline behind an auxiliary leg
            hamming_sharded = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (ADC scan): SURVEY 8d accounting
    cfg = ix.last_config()
    n_local = sh.hi - sh.lo
    qa, qb_ = sh.query_slice(nq)
    nq_rank = qb_ - qa                       # queries this rank's replica group answers
    row_bytes = ix.row_bytes
    lut_bytes = int(ix.lut_size) * 4
    n_launch = -(-nq_rank // cfg["queries_per_launch"])
    T = max(1, cfg["queries_per_cta"])
    # one pass over the local rows per query TILE (T queries share the stream) + per-query LUT and result bytes
    alg_bytes_step = -(-nq_rank // T) * n_local * row_bytes + nq_rank * (lut_bytes + k * 8 * cfg["row_chunks"])
    scan_ms_mean = float(np.mean(scan_ms))
    achieved = alg_bytes_step / (scan_ms_mean / 1e3) / 1e9
    kname = {1: "adc_scan_kernel", 2: "adc_filter_scan_kernel", 3: "adc_filter16_scan_kernel"}.get(cfg["scan_kernel"], "adc_scan_kernel")
    l2_resident = n_local * row_bytes <= 100e6 and nq_rank > T
    roofline = {"bound": "smem_lsu" if l2_resident else "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(kname, name, world), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes_step / n_launch, "launch_ms": scan_ms_mean / n_launch, "query_tile_T": T,
                "pairs_per_s": nq_rank * n_local / (scan_ms_mean / 1e3),
                "note": ("SURVEY 8d accounting: ceil(nq/T) passes over the packed rows (T queries share each pass; never multiplied "
                         "back by T) + LUT/result bytes, divided by the scan kernel's CUDA-event time. "
                         + (f"At this shape the packed codes of a GPU ({n_local * row_bytes / 1e6:.0f} MB) stay L2-resident while the query "
                            "tiles sweep them, so HBM is not the limiter (`traffic` = ncu DRAM bytes of one launch): the kernel is bound "
                            "by shared-memory table gathers (bound = smem_lsu); `peak` stays the HBM figure the accounting is defined "
                            "against.  roofline_hbm_shape has the same kernel on a shard >> L2." if l2_resident else
                            "The shard is far larger than L2: HBM-bound regime."))}

    # ---- LUT build: bytes written / FMA utilisation (north_star: "for the LUT build it is tensor-pipe or FMA utilisation")
    lut_ms_mean = float(np.mean(lut_ms))
    nq_first = min(nq_rank, cfg["queries_per_launch"])
    lut_written = nq_first * int(ix.lut_size) * (4 + (2 if cfg["scan_kernel"] == 3 else 0))
    lut_flops = nq_first * int(ix.lut_size) * 3 * model.L
    roofline_lut = {"bound": "hbm_write", "kernel": "lut_build_kernel", "launch_ms": lut_ms_mean, "bytes_written": lut_written,
                    "achieved": lut_written / (lut_ms_mean / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": lut_written / (lut_ms_mean / 1e3) / 1e9 / peak,
                    "flops": lut_flops, "fp32_tflops": lut_flops / (lut_ms_mean / 1e3) / 1e12,
                    "fma_utilisation": lut_flops / (lut_ms_mean / 1e3) / 1e12 / FP32_FMA_PEAK_TFLOPS,
                    "fp32_peak_tflops": FP32_FMA_PEAK_TFLOPS, "share_of_step": lut_ms_mean / (total_ms / args.steps),
                    "note": "first batch of the step: scale + table kernels; nq * sum(K) * 3L flop (sub, mul, add), fp32 + fp16 tables written"}

    # ---- the same scan kernel on a shard far larger than L2 (the HBM-bound regime of the 100M / 1B-row shapes)
    hbm_shape = None
    hamming = None
    if not args.no_hbm_shape and world == 1:
        try:
            big_n = args.hbm_rows
            big = VAQIndex(model.L, model.bits, model.centroids, device=local_rank)
            big.reserve(big_n)
            cdf = pb.cdf if pb.cdf is not None else synth.code_cdf(ix.get_codes(0, min(n_local, 200_000)), model.bits)
            big.add_synthetic(big_n, SEED, cdf)
            hbm_shape = {"kernel": "adc_filter16_scan_kernel", "rows": big_n, "packed_bytes": big_n * row_bytes, "peak": peak, "unit": "GB/s",
                         "bound": "hbm", "runs": []}
            for bq in (1, 8, 64):
                lab = torch.empty((bq, k), dtype=torch.int32, device=dev)
                dis = torch.empty((bq, k), dtype=torch.float32, device=dev)
                ms = []
                for i in range(3 + 5):
                    flush.fill_(i)
                    big.search_device(d_q.data_ptr(), bq, k, flags, lab.data_ptr(), dis.data_ptr(), st.cuda_stream)
                    torch.cuda.synchronize()
                    if i >= 3:
                        ms.append(big.last_timings()["scan_ms"])
                bcfg = big.last_config()
                bT = max(1, bcfg["queries_per_cta"])
                b = -(-bq // bT) * big_n * row_bytes + bq * lut_bytes
                a = b / (np.mean(ms) / 1e3) / 1e9
                hbm_shape["runs"].append({"queries": bq, "query_tile_T": bT, "scan_ms": float(np.mean(ms)), "achieved": a,
                                          "frac": a / peak, "traffic": ncu_traffic("adc_filter16_scan_kernel", f"hbm_shape_q{bq}", 1),
                                          "pairs_per_s": bq * big_n / (np.mean(ms) / 1e3), "config": bcfg})
            best = max(hbm_shape["runs"][:2], key=lambda r: r["frac"])       # one query tile = one pass over HBM
            hbm_shape.update({"achieved": best["achieved"], "frac": best["frac"], "traffic": best["traffic"]})
            big.close()
        except Exception as e:      # never hide the headline behind the auxiliary leg
            hbm_shape = {"error": repr(e)}
        try:
            hn = args.hbm_rows
            hx = HammingIndex(256, device=local_rank)
            hx.add_synthetic(hn, SEED)
            hamming = {"bound": "hbm", "kernel": "ham_scan_kernel", "rows": hn, "bits": 256, "packed_bytes": hn * 32, "peak": peak,
                       "unit": "GB/s", "runs": []}
            for hq in (1, 2, 8, 64):
                hqv = torch.from_numpy(synth.synth_bitvectors(hq, 10 ** 10, 256, SEED).view(np.int64)).to(dev)
                hidx = torch.empty((hq, k), dtype=torch.int32, device=dev)
                hdist = torch.empty((hq, k), dtype=torch.int32, device=dev)
                ms = []
                for i in range(3 + 5):
                    flush.fill_(i)
                    hx.query_device(hqv.data_ptr(), hq, k, hidx.data_ptr(), hdist.data_ptr(), st.cuda_stream)
                    torch.cuda.synchronize()
                    if i >= 3:
                        ms.append(hx.last_timings()["scan_ms"])
                hcfg = hx.last_config()
                qt = max(1, hcfg["queries_per_cta"])
                b = -(-hq // qt) * hn * 32
                a = b / (np.mean(ms) / 1e3) / 1e9
                hamming["runs"].append({"queries": hq, "queries_per_pass": qt, "scan_ms": float(np.mean(ms)), "achieved": a,
                                        "frac": a / peak, "traffic": ncu_traffic("ham_scan_kernel", f"ham_q{hq}", 1),
                                        "qps": hq / (np.mean(ms) / 1e3), "pairs_per_s": hq * hn / (np.mean(ms) / 1e3), "config": hcfg})
            best = max(hamming["runs"], key=lambda r: r["frac"])
            hamming.update({"achieved": best["achieved"], "frac": best["frac"], "traffic": best["traffic"],
                            "best_at_queries": best["queries"]})
            hx.close()
        except Exception as e:
            hamming = {"error": repr(e)}

    # ---- CPU baseline on the box's host cores + parity of the returned neighbours
    cpu = None
    parity = None
    if not args.no_cpu:
        from oracle import oracle as orc
        threads = os.cpu_count() or 1
        use_host_threads(threads)
        om = orc.Model(model.L, model.bits, model.centroids)
        glab_all, gdis_all = lab_pin.numpy(), dis_pin.numpy()
        if w["source"] == "vectors":
            parity = {}
            ti_ref = None
            if w.get("ti_clusters"):
                # the reference's own searchTriangleInequality on the clusters the device built: rows of a cluster sorted
                # far -> near the centre (VAQ.cpp:968-982), codeToCC by original id
                ti = ix.get_clusters()
                grouped = ix.get_codes()
                seg = w["ti_segments"]
                dec = np.concatenate([model.centroids[s_][grouped[:, s_]] for s_ in range(seg)], axis=1)
                cl_of = np.repeat(np.arange(ti["clusters"].shape[0]), ti["sizes"])
                c2c = np.sqrt(((dec - ti["clusters"][cl_of]) ** 2).sum(1)).astype(np.float32)
                order = np.lexsort((np.arange(c2c.size), -c2c, cl_of))
                c2c_by_id = np.empty_like(c2c)
                c2c_by_id[ti["members"][order]] = c2c[order]
                codes = grouped[order]
                ti_ref = dict(clusters=ti["clusters"], start=ti["start"], sizes=ti["sizes"], members=ti["members"][order],
                              code_to_cc=c2c_by_id, visit=w["visit"])
            else:
                codes, rows_used, how = reference_codes(pb, args.cpu_rows)      # VAQ::encode of the compiled reference, all rows
                if world == 1:
                    mine = ix.get_codes()
                    diff = np.nonzero((mine != codes).any(1))[0]
                    parity["device_encode_vs_reference_encode_mismatches"] = int((mine != codes).sum())
                    parity["codes_compared"] = int(codes.size)
                    if diff.size:       # float near-ties between two centroids (Eigen's reduction order)
                        _, margin = orc.Port().encode(om, pb.XP[diff], with_margin=True)
                        bad = mine[diff] != codes[diff]
                        parity["mismatch_margin_max"] = float(margin[bad].max())     # gap between the two centroids' distances
            qps, n_used, kind, rlab, rdis, dt = cpu_reference_search(model, codes, Q, k, args.cpu_seconds, threads, ti=ti_ref)
            cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind,
                   "sample": f"{n_used} of {nq} queries, all {n} rows, {'TI+EA mode on the device-built clusters' if ti_ref else 'EA mode'}, "
                             f"query-sliced over {threads} threads, {dt:.1f}s"}
            glab, gdis = glab_all[:n_used], gdis_all[:n_used]
            gt = synth.brute_force_knn(pb.X, pb.Qraw[:min(n_used, 200)], min(k, 10))
            kk = min(k, 10)
            parity.update({"checker": f"compiled {kind}: VAQ::search on the whole matrix", "queries": n_used,
                           "ids_equal_frac": float((glab == rlab).mean()),
                           "max_rel_dist_err": float(np.max(np.abs(gdis - rdis) / np.maximum(rdis, 1e-30))),
                           "recall_at_10_gpu": synth.recall_at_k(glab[:gt.shape[0]], gt, kk),
                           "recall_at_10_reference": synth.recall_at_k(rlab[:gt.shape[0]], gt, kk)})
        else:
            # the host cannot hold the matrix: CPU timing on a row slice (scaled), parity through size-independent checks
            codes, rows_used, how = reference_codes(pb, args.cpu_rows)
            qps, n_used, kind, rlab, rdis, dt = cpu_reference_search(model, codes, Q, k, args.cpu_seconds, threads)
            cpu = {"value": qps * rows_used / n, "unit": "queries/s", "cores": threads, "kind": kind,
                   "sample": f"{n_used} of {nq} queries, {how}, EA mode, query-sliced over {threads} threads, {dt:.1f}s"}
            port = orc.Port()
            nchk = min(nq, 16)
            lut = port.create_lut(om, Q[:nchk])
            ok_dist = ok_slice = ok_sorted = True
            for q in range(nchk):
                got = np.concatenate([pb.host_codes(int(r), 1) for r in glab_all[q]])
                ok_dist &= bool(np.array_equal(port.adc_all(om, lut[q], got).view(np.uint32), gdis_all[q].view(np.uint32)))
                d = port.adc_all(om, lut[q], codes)
                better = np.nonzero(d < gdis_all[q, -1])[0]
                ok_slice &= set(better.tolist()) <= set(glab_all[q].tolist())
                ok_sorted &= bool((np.diff(gdis_all[q]) >= 0).all())
            parity = {"checker": "oracle port on host-regenerated rows", "queries": nchk,
                      "returned_distances_bit_equal_oracle": bool(ok_dist), "no_better_row_in_host_slice": bool(ok_slice),
                      "sorted": bool(ok_sorted), "host_slice_rows": rows_used}
            nl, nd = needle_out
            lutn = port.create_lut(om, needle_q)
            parity["planted_needles_ok"] = bool(all(
                nd[j, 0] == port.adc_all(om, lutn[j], pb.host_codes(needles[j], 1))[0] and (nl[j, 0] == needles[j] or nd[j, 0] == nd[j, 1])
                for j in range(len(needles))))
            parity["planted_needles"] = len(needles)

    line = {
        "metric": "queries/sec at recall@10", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(name, w, "TI+EA, visit %.2f of %d clusters" % (w["visit"], w["ti_clusters"]) if w.get("ti_clusters") else "EA"),
                       row_bytes=row_bytes,
                       sharding=f"rows/{sh.R}" + (f" x queries/{sh.qgroups} (matrix replicated {sh.qgroups}x)" if sh.qgroups > 1 else ""),
                       rows_per_gpu=n_local, bound_exchange="nvlink peer memory" if exchange else "off",
                       l2="256 MB fill between timed steps", source=w["source"], desc=w["desc"], scan_config=cfg),
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(Q.nbytes), "d2h_bytes_per_step": int(nq * k * 8)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "roofline_lut_build": roofline_lut,
        "roofline_hbm_shape": hbm_shape,
        "roofline_hamming": hamming,
        "hamming_sharded": hamming_sharded,
        "cpu_baseline": cpu,
        "parity_vs_cpu": parity,
        "kernel_ms": {"lut_build": lut_ms_mean, "adc_scan": scan_ms_mean, "merge": float(np.mean(merge_ms)),
                      "max_over_ranks": {"adc_scan": scan_max[0], "lut_build": scan_max[1], "merge": scan_max[2]}},
        "step_ms_rank0": [float(x) for x in step_ms], "step_scan_ms_rank0": [float(x) for x in scan_ms],
    }
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sift1m_256b_m32_k10", choices=sorted(WORKLOADS))
    ap.add_argument("--queries", type=int, default=0, help="override the workload's query count")
    ap.add_argument("--rows", type=int, default=0, help="override the workload's row count (rehearsals; the result is not the named config)")
    ap.add_argument("--visit", type=float, default=0.0, help="override the TI workload's visit fraction")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-rows", type=int, default=2_000_000, help="host row slice for the CPU baseline of device-generated workloads")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hbm-shape", action="store_true")
    ap.add_argument("--no-hamming", action="store_true")
    ap.add_argument("--no-bound-exchange", action="store_true", help="comparison: shards prune with their own bounds only")
    ap.add_argument("--hbm-rows", type=int, default=64_000_000)
    ap.add_argument("--ham-rows", type=int, default=256_000_000, help="total rows of the row-sharded Hamming leg")
    ap.add_argument("--ham-queries", type=int, default=64)
    ap.add_argument("--scan-v1", action="store_true", help="force the lane-per-row scan kernel (comparison)")
    ap.add_argument("--row-shards", type=int, default=0,
                    help="experiment knob; default 0 = N row shards (pure row sharding, the BASELINE layout).  R < N replicates "
                         "the code matrix N/R times and splits the query batch between the replica groups")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = dict(WORKLOADS[args.workload])
    if args.queries:
        w["nq"] = args.queries
    if args.rows:
        w["n"] = args.rows
        w["desc"] += f" [rows overridden: {args.rows}]"
        if "ham_rows" in w:
            w["ham_rows"] = args.rows
    if args.visit and "visit" in w:
        w["visit"] = args.visit
    if args.impl == "reference":
        run_reference_arm(args, w, args.workload)
    else:
        run_gpu_arm(args, w, args.workload)


if __name__ == "__main__":
    main()
