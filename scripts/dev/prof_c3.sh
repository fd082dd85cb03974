# ncu evidence for the spill path (C3): adc_filter_scan_kernel<4,1> and merge_level_kernel
mkdir -p gpurun_out
CMD="python bench.py --workload gist1m_512b_m64_k10 --steps 1 --warmup 3 --no-cpu --no-hbm-shape --no-hamming"
timeout 200 $CMD > gpurun_out/r2z_bench_c3.json 2> gpurun_out/r2z_bench_c3.log || { echo "plain bench failed"; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"adc_filter_scan_kernel|merge_level" -s 4 -c 2 -f -o gpurun_out/r2z_c3 $CMD > gpurun_out/r2z_ncu_c3.log 2>&1; echo "ncu rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2z_bench_c3.json')); print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['roofline']['frac'], d['config']['scan_config'])"
