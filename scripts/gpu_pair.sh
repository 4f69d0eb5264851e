( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -p no:cacheprovider -k "edgeconv" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head -20
