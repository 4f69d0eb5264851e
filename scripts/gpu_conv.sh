mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -k "conv_bf16" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -20
echo "== multicast"; python scripts/kbench_conv.py 2>&1 | grep -v Warning | tee gpurun_out/kbench_conv.txt
echo "== no multicast"; CP_CONV_MULTICAST=0 python scripts/kbench_conv.py 2>&1 | grep -v Warning
