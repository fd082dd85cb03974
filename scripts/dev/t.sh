t() { echo "== $1 [$2]"; VAQGPU_LIB="$1" VAQGPU_TUNE="$2" timeout 60 python -m pytest tests/test_gpu_vaq.py -x -q -m gpu -k "multi_chunk" 2>&1 | tail -1; }
t "" ""
t $PWD/vaq_b200/libvaqgpu_oldexact.so ""
t "" "order=0"
t "" "rot=0"
t "" "seed=0"
t "" "q3cap=1"
