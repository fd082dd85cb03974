run() { # name, tune, workload, [lib]
  echo "== $1 [$2] $3 $4"
  VAQGPU_LIB="$4" VAQGPU_TUNE="$2" timeout 120 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu --no-hbm-shape --no-hamming 2>gpurun_out/r2t_$1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['kernel_ms'].get('adc_scan'), d['kernel_ms'].get('lut_build'), d['config']['scan_config']['row_chunks'], d['config']['scan_config']['layout_us'], d['ms_per_step'])"
  grep "dbg\|stats" gpurun_out/r2t_$1.log | tail -2
}
timeout 600 python -m pytest tests/test_gpu_vaq.py -x -q -m gpu 2>&1 | tail -3
S=$PWD/vaq_b200/libvaqgpu_stats.so
run s125 "" shard125k_256b_m32_k10
run s125_noorder "order=0" shard125k_256b_m32_k10
run s125_keep "keepthr=1" shard125k_256b_m32_k10
run s125_stats "dbg=1" shard125k_256b_m32_k10 $S
run s125_spl8 "spl=8" shard125k_256b_m32_k10
run s125_q3cap4 "q3cap=4" shard125k_256b_m32_k10
run s1m "" sift1m_256b_m32_k10
run s1m_noorder "order=0" sift1m_256b_m32_k10
run s1m_keep "keepthr=1" sift1m_256b_m32_k10
run s1m_stats "dbg=1" sift1m_256b_m32_k10 $S
for t in ""; do VAQGPU_TUNE="$t" timeout 200 python scripts/dev/conflicts.py shard125k_256b_m32_k10 2>/dev/null | tail -1; done
for t in ""; do VAQGPU_TUNE="$t" timeout 200 python scripts/dev/conflicts.py sift1m_256b_m32_k10 2>/dev/null | tail -1; done
