run() { # name, tune, workload, [lib]
  echo "== $1 [$2] $3 $4"
  VAQGPU_LIB="$4" VAQGPU_TUNE="$2" timeout 120 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu --no-hbm-shape --no-hamming 2>gpurun_out/r2s_$1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['kernel_ms'].get('adc_scan'), d['kernel_ms'].get('lut_build'), d['config']['scan_config']['row_chunks'], d['config']['scan_config']['layout_us'], d['ms_per_step'])"
  grep "dbg\|stats" gpurun_out/r2s_$1.log | tail -2
}
S=$PWD/vaq_b200/libvaqgpu_stats.so
run s125_stats "dbg=1" shard125k_256b_m32_k10 $S
run s125_stats_noorder "dbg=1,order=0" shard125k_256b_m32_k10 $S
run s125_stats_keep "dbg=1,keepthr=1" shard125k_256b_m32_k10 $S
run s1m_stats "dbg=1" sift1m_256b_m32_k10 $S
run s1m_stats_noorder "dbg=1,order=0" sift1m_256b_m32_k10 $S
run s1m_stats_keep "dbg=1,keepthr=1" sift1m_256b_m32_k10 $S
run s1m_stats_keep_noorder "dbg=1,keepthr=1,order=0" sift1m_256b_m32_k10 $S
run s1m_keep_noorder "keepthr=1,order=0" sift1m_256b_m32_k10
run s1m_keep_norot "keepthr=1,rot=0" sift1m_256b_m32_k10
run s1m_norot "rot=0" sift1m_256b_m32_k10
