mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -k "slab or border" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -20
timeout 300 python scripts/kbench_slab.py 2>&1 | grep -E "^conv" | tee gpurun_out/kbench_slab.txt
echo "== ring"; KB_SHAPES=1 CP_SLAB_RESIDENT=0 timeout 300 python scripts/kbench_slab.py 2>&1 | grep -E "^conv"
