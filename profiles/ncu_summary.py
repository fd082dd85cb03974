#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest source lines.  usage: ncu_summary.py rep [nlines]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__cycles_elapsed.avg', 'smsp__inst_executed_op_shared_ld.sum',
        'l1tex__data_pipe_lsu_wavefronts.sum', 'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_barrier_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct',
        'smsp__warp_issue_stalled_not_selected_per_warp_active.pct', 'smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct']
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")][:60] if "Kernel Name" in hdr else "")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:80s} {r[i]:>20s} {units[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[0]
    def col(name):
        for i, x in enumerate(h):
            if x.strip() == name:
                return i
        return None
    ci = col("Instructions Executed"); cs = col("Warp Stall Sampling (All Samples)") or col("Warp Stall Sampling (All Cycles)") ; csrc = col("Source")
    print("columns:", [x for x in h][:12])
    data = []
    for r in rows[1:]:
        try:
            data.append((float(r[cs] or 0), float(r[ci] or 0), r[csrc] if csrc is not None else ""))
        except Exception:
            pass
    tot_s = sum(d[0] for d in data) or 1; tot_i = sum(d[1] for d in data) or 1
    print(f"-- hottest source lines by stall samples (total samples {tot_s:.0f}, total inst {tot_i:.0f})")
    for d in sorted(data, key=lambda x: -x[0])[:nl]:
        print(f"{100*d[0]/tot_s:6.2f}% samples {100*d[1]/tot_i:6.2f}% inst | {d[2].strip()[:130]}")
