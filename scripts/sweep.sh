for t in "seed=1" "seed=0" "seed=1,chunks=2" "seed=1,chunks=4" "seed=1,threads=512" "seed=1,T=2" ; do
  echo "== $t"; VAQGPU_TUNE="$t" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-hbm-shape 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), d['kernel_ms'], d['config']['scan_config'])"
done
