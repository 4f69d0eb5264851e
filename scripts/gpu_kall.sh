# kernel parity tests + micro-benchmark of all three fused kernels for the in-tree build
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -p no:cacheprovider 2>&1 | tail -3
python scripts/kbench.py 2>&1 | grep -v Warning
