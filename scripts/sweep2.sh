for w in shard125k_256b_m32_k10 tiny16k_256b_m32_k10 sift1m_256b_m32_k10; do for t in "seed=1"; do
  echo "== $w $t"; VAQGPU_TUNE="$t" timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-hbm-shape 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['kernel_ms'], d['config']['scan_config']['threads'])"
done; done
