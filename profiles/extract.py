#!/usr/bin/env python
"""Turn .ncu-rep captures (gpurun_out/, scratch) into the small text artefacts committed here.
usage: [NCU_ROW=i] python profiles/extract.py <tag> <rep> [<kernel-key> <workload> <n_gpus>]"""
import csv, io, json, subprocess, sys
from pathlib import Path
HERE = Path(__file__).resolve().parent
tag, rep = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
import os
hdr, units, vals = rows[0], rows[1], rows[2 + int(os.environ.get("NCU_ROW", "0"))]      # NCU_ROW: which result of a multi-kernel report
keep = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "launch__shared_mem_per_block_dynamic"]
out = {}
for k in keep:
    if k in hdr:
        i = hdr.index(k)
        out[k] = f"{vals[i]} {units[i]}".strip()
stalls = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): int(float(vals[i])) for i, h in enumerate(hdr)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("not_issued")}
tot = sum(stalls.values()) or 1
out["stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -x[1])[:8]}
(HERE / f"{tag}_summary.json").write_text(json.dumps(out, indent=1))
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True).stdout
(HERE / f"{tag}_details.csv").write_text(det)
print(json.dumps(out, indent=1))
if len(sys.argv) > 5:
    def num(k):
        v, u = out[k].split()[0], out[k].split()[1] if len(out[k].split()) > 1 else ""
        m = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        return float(v) * m
    tj_path = HERE / "ncu_traffic.json"
    tj = json.loads(tj_path.read_text()) if tj_path.exists() else {}
    ent = {"workload": sys.argv[4], "n_gpus": int(sys.argv[5]), "dram_bytes_read": num("dram__bytes_read.sum"),
           "dram_bytes_write": num("dram__bytes_write.sum"), "gpu_time": out.get("gpu__time_duration.sum"),
           "source": f"profiles/{tag}_summary.json"}
    cur = tj.get(sys.argv[3], [])
    if isinstance(cur, dict):
        cur = [cur]
    cur = [e for e in cur if not (e.get("workload") == ent["workload"] and e.get("n_gpus") == ent["n_gpus"])] + [ent]
    tj[sys.argv[3]] = cur
    tj_path.write_text(json.dumps(tj, indent=1))
