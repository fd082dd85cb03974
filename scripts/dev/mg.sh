n=$1
r() { echo "== $1 [$2]"; env $2 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 5 --warmup 3 --no-cpu --no-hamming 2>gpurun_out/r2y_$1.log | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), d['ms_per_step'], d['config']['bound_exchange'], d['step_ms_rank0'], d['step_scan_ms_rank0'], d['clocks'])"; }
r a1 "A=1"
r a2 "A=1"
r ns1 "VAQ_BENCH_NO_SAMPLER=1"
r ns2 "VAQ_BENCH_NO_SAMPLER=1"
