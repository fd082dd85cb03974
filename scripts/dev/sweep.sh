timeout 300 python -m pytest tests/test_gpu_hamming.py -x -q -m gpu 2>&1 | tail -2
for q in 1 2; do echo "== nq=$q"; timeout 120 python scripts/ham_bench.py 64000000 $q 2>&1 | tail -1; done
echo "== nq=2 hamocc=4"; VAQGPU_TUNE=hamocc=4 timeout 120 python scripts/ham_bench.py 64000000 2 2>&1 | tail -1
