run() { # name, tune, workload, [lib]
  echo "== $1 [$2] $3 $4"
  VAQGPU_LIB="$4" VAQGPU_TUNE="$2" timeout 120 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu --no-hbm-shape --no-hamming 2>gpurun_out/r2u_$1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['kernel_ms'].get('adc_scan'), d['kernel_ms'].get('lut_build'), d['config']['scan_config']['row_chunks'], d['config']['scan_config']['layout_us'], d['ms_per_step'])"
  grep "dbg\|stats" gpurun_out/r2u_$1.log | tail -2
}
for c in 16 32 128; do run s125_c$c "order_c=$c" shard125k_256b_m32_k10; done
for c in 32 64 128 512; do run s1m_c$c "order_c=$c" sift1m_256b_m32_k10; done
run c1 "" siftsmall_128b_m16_k100
run small "" small_256b_m32_k10
