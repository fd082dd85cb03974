run() { # name, tune, workload, [lib]
  echo "== $1 [$2] $3 $4"
  VAQGPU_LIB="$4" VAQGPU_TUNE="$2" timeout 200 python bench.py --workload $3 --steps 3 --warmup 3 --no-cpu --no-hbm-shape --no-hamming 2>gpurun_out/r2ag_$1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['kernel_ms'].get('adc_scan'), d['kernel_ms'].get('lut_build'), d['ms_per_step'])"
}
timeout 600 python -m pytest tests/test_gpu_vaq.py tests/test_gpu_api.py -x -q -m gpu 2>&1 | tail -2
run s125_tpc4 "" shard125k_256b_m32_k10
run s125_tpc1 "luttpc=1" shard125k_256b_m32_k10
run s125_tpc2 "luttpc=2" shard125k_256b_m32_k10
run s125_tpc8 "luttpc=8" shard125k_256b_m32_k10
