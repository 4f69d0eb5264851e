mkdir -p gpurun_out
KB_CHECK_ONLY=1 timeout 120 python scripts/kbench_slab.py 2>&1 | grep -v Warning
echo "== pair"; KB_SHAPES=1 timeout 300 python scripts/kbench_slab.py 2>&1 | grep -E "^conv"
