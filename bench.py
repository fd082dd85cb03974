#!/usr/bin/env python
"""Headline benchmark: VAQ query-time search (LUT build -> ADC scan -> top-k) on B200.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU search

Workload (BASELINE.json configs[1]): SIFT1M-shape synthetic, 1M x 128 base, 10K queries,
VAQ 256-bit budget over 32 subspaces, k = 10.  A step = one search of the whole 10K-query batch
against the index.  N > 1 (torchrun, one rank per GPU): the code matrix is row-sharded, every
rank scans its rows for all queries, one NCCL all-gather of the shard-local top-k + device merge
(strong scaling: total rows fixed).

One JSON line on stdout (rank 0); everything else goes to stderr.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: rows, dims, budget, M, min_bits, max_bits, queries, k, decay
    "sift1m_256b_m32_k10": dict(n=1_000_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=10_000, k=10, decay=4.0),
    "small_256b_m32_k10": dict(n=100_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=1_000, k=10, decay=4.0),
    "shard125k_256b_m32_k10": dict(n=125_000, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=10_000, k=10, decay=4.0),
    "tiny16k_256b_m32_k10": dict(n=16_384, d=128, budget=256, M=32, min_bits=7, max_bits=9, nq=10_000, k=10, decay=4.0),
}
TRAIN_ROWS = 32768
SEED = 13517106


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


def build_problem(w: dict):
    """Seeded synthetic base/queries + host-trained model (the reference keeps training on the host)."""
    from vaq_b200 import synth, train
    t0 = time.time()
    X = synth.decaying_gaussian(w["n"], w["d"], decay=w["decay"], seed=SEED)
    Qraw = synth.decaying_gaussian(w["nq"], w["d"], decay=w["decay"], seed=SEED + 7)
    model, _ = train.train(X[:TRAIN_ROWS], w["budget"], w["M"], w["min_bits"], w["max_bits"], kmeans_iters=8, seed=SEED)
    XP = model.project(X)
    Q = model.project(Qraw)
    log(f"[bench] problem built in {time.time() - t0:.1f}s bits={model.bits.tolist()}")
    return model, X, XP, Qraw, Q


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


# ---------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU search (compiled unmodified reference when available)
# ---------------------------------------------------------------------------------------------

def cpu_reference_search(model, codes, Q, k, seconds_target: float, threads: int):
    """Times reference VAQ::search (EA mode, the reference's fastest exact mode) on a bounded sample of
    the query batch.  Returns (qps, n_queries, kind, labels, dists)."""
    from oracle import oracle as orc
    om = orc.Model(model.L, model.bits, model.centroids)
    if orc.Ref.available():
        kind = "reference"
        rv = orc.Ref().vaq(om, orc.NN_EA)
        rv.set_codes(codes)

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


def run_reference_arm(args, w, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as orc
    model, X, XP, Qraw, Q = build_problem(w)
    om = orc.Model(model.L, model.bits, model.centroids)
    t0 = time.time()
    if orc.Ref.available():      # VAQ::encode of the compiled reference (OpenMP over rows, VAQ.cpp:733)
        rv = orc.Ref().vaq(om, orc.NN_EA)
        codes = rv.encode(XP)
        rv.close()
    else:
        codes = orc.Port().encode(om, XP)
    log(f"[bench] host encode {time.time() - t0:.1f}s")
    threads = os.cpu_count() or 1
    vals = []
    n_used = 0
    kind = "reference"
    per_step_target = max(2.0, min(20.0, 60.0 / max(1, args.steps + args.warmup)))
    for it in range(args.warmup + args.steps):
        qps, n_used, kind, _, _, dt = cpu_reference_search(model, codes, Q, w["k"], per_step_target, threads)
        if it >= args.warmup:
            vals.append((qps, dt))
    qps = float(np.mean([v[0] for v in vals]))
    ms = float(np.mean([v[1] for v in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": "queries/sec at recall@10", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "rows": w["n"], "dims": w["d"], "queries": w["nq"], "k": w["k"], "bits": w["budget"],
                   "subspaces": w["M"], "mode": "EA (VAQ::searchEarlyAbandon)"},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind,
                         "sample": f"{n_used} of {w['nq']} queries per step, all {w['n']} rows, query-sliced over {threads} threads"},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------------------
# this repo's arm
# ---------------------------------------------------------------------------------------------

def run_gpu_arm(args, w, name):
    import torch
    import torch.distributed as dist
    from vaq_b200 import synth
    from vaq_b200.index import EA, PROJECTED, VAQIndex
    from vaq_b200.sharded import ShardedVAQ, shard_bounds

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
        os.environ["NCCL_DEBUG"] = "WARN"          # keep NCCL's version banner off stdout (one JSON line only)
        dist.init_process_group("nccl", device_id=dev)

    model, X, XP, Qraw, Q = build_problem(w)
    n, nq, k = w["n"], w["nq"], w["k"]
    flags = EA | PROJECTED | (0x1000 if args.scan_v1 else 0)

    # index: this rank's row block, encoded on the device (bit-exact vs the oracle, tests/test_gpu_vaq.py)
    # Layout.  --row-shards R: R ranks share one copy of the code matrix (row-sharded), world/R replica groups split
    # the query batch.  Default ("auto"): the fewest row shards that keep a GPU's share of the packed matrix under
    # 16 GiB — a 32 MB matrix is replicated, a 1B-row one is row-sharded.  The pure row-sharded layout BASELINE.json
    # names (R = N) is always timed as well (`row_sharded_layout` in the JSON line).
    packed_bytes = n * 16 * (-(-w["budget"] // 128))
    R = args.row_shards
    if not R:
        R = 1
        while R < world and packed_bytes / R > (16 << 30):
            R *= 2
    sh = ShardedVAQ(model.L, model.bits, model.centroids, model.eig, n, rank, world, local_rank, row_shards=R)
    t0 = time.time()
    sh.index.encode_add(XP[sh.lo:sh.hi])
    log(f"[bench] rank {rank}: encoded rows [{sh.lo},{sh.hi}) in {time.time() - t0:.1f}s; row_bytes={sh.index.row_bytes}")
    ix = sh.index

    sh_rows = None
    if world > 1 and sh.R < world:
        sh_rows = ShardedVAQ(model.L, model.bits, model.centroids, model.eig, n, rank, world, local_rank, row_shards=world)
        sh_rows.index.encode_add(XP[sh_rows.lo:sh_rows.hi])

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
    sampler = ClockSampler(local_rank) if rank == 0 else None      # started before warm-up: NVML start-up stays out of the timed steps
    for _ in range(args.warmup):
        step_device()
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
        launches += ix.last_config()["launches"] + (1 if world > 1 else 0) + 1   # + all-gather + final merge
    barrier()
    clocks = sampler.stop() if sampler else None
    step_ms = np.array([a.elapsed_time(b) for a, b in ev], dtype=np.float64)
    total_ms = torch.tensor([step_ms.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = nq * args.steps / (total_ms / 1e3)

    # ---- the same steps on the pure row-sharded layout (N shards, all queries on every rank), when it is not the default
    row_layout = None
    if sh_rows is not None:
        for _ in range(args.warmup):
            sh_rows.search(d_q, k, flags)
        barrier()
        ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(i & 0xFF)
            barrier()
            ev2[i][0].record(st)
            rl, rd = sh_rows.search(d_q, k, flags)
            ev2[i][1].record(st)
            torch.cuda.synchronize()
        barrier()
        t2 = torch.tensor([sum(a.elapsed_time(b) for a, b in ev2)], dtype=torch.float64, device=dev)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        same = bool(torch.equal(rl, labels) and torch.equal(rd.view(torch.int32), dists.view(torch.int32)))
        row_layout = {"sharding": f"rows/{world}", "value": nq * args.steps / (float(t2.item()) / 1e3), "unit": "queries/s",
                      "ms_per_step": float(t2.item()) / args.steps, "identical_to_default_layout": same,
                      "scan_ms": sh_rows.index.last_timings()["scan_ms"]}

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

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (ADC scan), T = 1 accounting (each query's CTAs stream the rows)
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
    peak, peak_src = measured_peak_hbm()
    achieved = alg_bytes_step / (scan_ms_mean / 1e3) / 1e9
    kname = {1: "adc_scan_kernel", 2: "adc_filter_scan_kernel", 3: "adc_filter16_scan_kernel"}.get(cfg["scan_kernel"], "adc_scan_kernel")
    traffic = None
    try:        # DRAM bytes per launch of this kernel on this workload from the committed ncu --set full capture
        tj = json.loads((ROOT / "profiles" / "ncu_traffic.json").read_text())
        ent = tj.get(kname, {})
        if ent.get("workload") == name and ent.get("n_gpus") == world:
            traffic = ent.get("dram_bytes_read", 0) + ent.get("dram_bytes_write", 0)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_step / n_launch,
                "launch_ms": scan_ms_mean / n_launch, "query_tile_T": T,
                "pairs_per_s": nq_rank * n_local / (scan_ms_mean / 1e3),
                "note": ("SURVEY 8d accounting: ceil(nq/T) passes over the packed rows (T queries share each pass) + LUT/result bytes; "
                         f"at this shape the packed codes ({n_local * row_bytes / 1e6:.0f} MB) are L2-resident and the scan is bound by "
                         "shared-memory LUT gathers (ncu: LSU pipe 87 %, issue 62 %, DRAM 1.5 %), not HBM; `traffic` = ncu DRAM bytes "
                         "of one launch — see roofline_hbm_shape for the same kernel on a shard >> L2")}

    # ---- the same kernel on a shard far larger than L2 (the HBM-bound regime of the 100M / 1B-row shapes)
    hbm_shape = None
    hamming = None
    if not args.no_hbm_shape:
        try:
            big_n = args.hbm_rows
            big = VAQIndex(model.L, model.bits, model.centroids, device=local_rank)
            big.reserve(big_n)
            cdf = synth.code_cdf(ix.get_codes(0, min(n_local, 200_000)), model.bits)
            big.add_synthetic(big_n, SEED, cdf)
            hbm_shape = {"rows": big_n, "packed_bytes": big_n * row_bytes, "peak": peak, "unit": "GB/s", "runs": []}
            for bq in (4, 64):
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
                                          "frac": a / peak, "first_word_stream_GBps": -(-bq // bT) * big_n * 16 / (np.mean(ms) / 1e3) / 1e9,
                                          "pairs_per_s": bq * big_n / (np.mean(ms) / 1e3), "config": bcfg})
            big.close()
        except Exception as e:      # never hide the headline behind the auxiliary leg
            hbm_shape = {"error": repr(e)}
        try:
            from vaq_b200.index import HammingIndex
            hn = args.hbm_rows
            hx = HammingIndex(256, device=local_rank)
            hx.add_synthetic(hn, SEED)
            hamming = {"kernel": "ham_scan_kernel", "rows": hn, "bits": 256, "packed_bytes": hn * 32, "peak": peak, "unit": "GB/s",
                       "runs": []}
            for hq in (2, 64):
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
                                        "frac": a / peak, "qps": hq / (np.mean(ms) / 1e3),
                                        "pairs_per_s": hq * hn / (np.mean(ms) / 1e3), "config": hcfg})
            hx.close()
        except Exception as e:
            hamming = {"error": repr(e)}

    # ---- CPU baseline on the box's host cores + parity of the returned neighbours on the same queries
    cpu = None
    parity = None
    if not args.no_cpu and world > 1:
        # multi-GPU: the merged answer of the first queries against the oracle's canonical answer on the whole matrix
        from oracle import oracle as orc
        om = orc.Model(model.L, model.bits, model.centroids)
        nchk = 32
        try:        # torchrun exports OMP_NUM_THREADS=1; the checker may use the idle host cores
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(max(1, (os.cpu_count() or 1) // 2))
        except OSError:
            pass
        codes_all = orc.Port().encode(om, XP)
        wl, wd = orc.Port().search_lex(om, codes_all, Q[:nchk], k)
        glab, gdis = lab_pin.numpy()[:nchk], dis_pin.numpy()[:nchk]
        parity = {"queries": nchk, "checker": "oracle port, whole matrix", "ids_equal_frac": float((glab == wl).mean()),
                  "dists_bit_equal": bool(np.array_equal(gdis.view(np.uint32), wd.view(np.uint32)))}
    if not args.no_cpu and world == 1:
        codes_local = ix.get_codes()
        if world == 1:
            threads = os.cpu_count() or 1
            qps, n_used, kind, rlab, rdis, dt = cpu_reference_search(model, codes_local, Q, k, args.cpu_seconds, threads)
            cpu = {"value": qps, "unit": "queries/s", "cores": threads, "kind": kind,
                   "sample": f"{n_used} of {nq} queries, all {n} rows, EA mode, query-sliced over {threads} threads, {dt:.1f}s"}
            glab = lab_pin.numpy()[:n_used]
            gdis = dis_pin.numpy()[:n_used]
            same = float((glab == rlab).mean())
            rel = float(np.max(np.abs(gdis - rdis) / np.maximum(rdis, 1e-30)))
            gt = synth.brute_force_knn(X, Qraw[:min(n_used, 200)], k)
            parity = {"queries": n_used, "ids_equal_frac": same, "max_rel_dist_err": rel,
                      "recall_at_10_gpu": synth.recall_at_k(glab[:gt.shape[0]], gt, k),
                      "recall_at_10_reference": synth.recall_at_k(rlab[:gt.shape[0]], gt, k)}

    line = {
        "metric": "queries/sec at recall@10", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "rows": n, "dims": w["d"], "queries": nq, "k": k, "bits": w["budget"], "subspaces": w["M"],
                   "mode": "EA", "row_bytes": row_bytes,
                   "sharding": f"rows/{sh.R}" + (f" x queries/{sh.qgroups} (matrix replicated {sh.qgroups}x)" if sh.qgroups > 1 else ""), "l2": "256 MB fill between timed steps",
                   "scan_config": cfg},
        "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": int(Q.nbytes), "d2h_bytes_per_step": int(nq * k * 8)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": roofline,
        "row_sharded_layout": row_layout,
        "roofline_hbm_shape": hbm_shape,
        "hamming_scan": hamming,
        "cpu_baseline": cpu,
        "parity_vs_cpu": parity,
        "kernel_ms": {"lut_build": float(np.mean(lut_ms)), "adc_scan": scan_ms_mean, "merge": float(np.mean(merge_ms))},
        "step_ms_rank0": [float(x) for x in step_ms],
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
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-hbm-shape", action="store_true")
    ap.add_argument("--hbm-rows", type=int, default=64_000_000)
    ap.add_argument("--scan-v1", action="store_true", help="force the lane-per-row scan kernel (comparison)")
    ap.add_argument("--row-shards", type=int, default=0,
                    help="R: ranks per replica group (default N = pure row sharding, the BASELINE layout); R < N replicates the "
                         "code matrix N/R times and splits the query batch between the groups")
    args = ap.parse_args()
    capture_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w, args.workload)
    else:
        run_gpu_arm(args, w, args.workload)


if __name__ == "__main__":
    main()
