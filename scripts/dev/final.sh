# end-of-round validation on one GPU: GPU tests, smoke(), the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2z_bench_1gpu.json 2> gpurun_out/r2z_bench_1gpu.log; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2z_bench_ref.json 2> gpurun_out/r2z_bench_ref.log; echo "ref rc=$?"
python - <<'P'
import json
d=json.load(open('gpurun_out/r2z_bench_1gpu.json'))
print(round(d['value']), d['ms_per_step'], d['kernel_ms'], d['e2e'], d['gpu_launches'], d['clocks'])
print('roofline', {k:d['roofline'][k] for k in ('bound','achieved','frac','pairs_per_s','traffic')})
print('hbm', [(r['queries'], round(r['frac'],3)) for r in d['roofline_hbm_shape']['runs']])
print('ham', [(r['queries'], round(r['frac'],3)) for r in d['roofline_hamming']['runs']])
print('cpu', d.get('cpu_baseline'))
print('parity', d.get('parity_vs_cpu'))
r=json.load(open('gpurun_out/r2z_bench_ref.json')); print('ref', r.get('value'), r.get('cpu_baseline'))
P
