r() { n=$1; w=$2; shift 2; echo "== N=$n $w $*"; timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --steps 5 --warmup 3 "$@" 2>gpurun_out/r2ae_${n}_$w.log > gpurun_out/r2ae_${n}_$w.json; python - <<P
import json
d=json.load(open('gpurun_out/r2ae_${n}_$w.json'))
print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['kernel_ms']['max_over_ranks'], d['config']['bound_exchange'], [round(x,3) for x in d['step_ms_rank0']], 'e2e', round(d['e2e']['value']))
print('   parity', d.get('parity_vs_cpu'))
P
}
timeout 400 python -m pytest tests -x -q -m gpu 2>&1 | tail -2
r 8 sift1m_256b_m32_k10 --no-hamming --cpu-seconds 4
r 8 sift1m_256b_m32_k10 --no-hamming --no-cpu
r 4 sift1m_256b_m32_k10 --no-hamming --no-cpu
r 2 sift1m_256b_m32_k10 --no-hamming --no-cpu
r 8 deep100m_128b_m16_k10 --no-hamming --cpu-seconds 3
