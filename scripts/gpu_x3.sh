mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -k "x3" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -30
