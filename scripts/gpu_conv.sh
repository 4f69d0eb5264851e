mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -s -k "conv_bf16" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED|bf16 conv" | head -20
python scripts/kbench_conv.py 2>&1 | grep -v Warning | tee gpurun_out/kbench_conv.txt
