# ncu evidence for profiles/: launch list of the bench command + full captures of the three kernels of a search
mkdir -p gpurun_out
timeout 300 python scripts/prof_run.py sift1m 4 > gpurun_out/r2z_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-hbm-shape --no-hamming > gpurun_out/r2z_bench_plain.json 2> gpurun_out/r2z_bench_plain.log || { echo "plain bench failed"; exit 1; }
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2z_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-hbm-shape --no-hamming > gpurun_out/r2z_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"adc_filter16|lut_build_kernel|merge_level" -s 6 -c 3 -f -o gpurun_out/r2z_search_sift1m python scripts/prof_run.py sift1m 4 > gpurun_out/r2z_ncu_full.log 2>&1; echo "full rc=$?"
ls -la gpurun_out/r2z_*
