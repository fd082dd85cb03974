for w in tiny16k_256b_m32_k10 shard125k_256b_m32_k10 sift1m_256b_m32_k10; do
  echo "== $w"; VAQGPU_TUNE="dbg=1" timeout 300 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu --no-hbm-shape 2>&1 | grep "vaqgpu dbg" | tail -1
done
