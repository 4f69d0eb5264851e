mkdir -p gpurun_out
CP_EDGECONV_PAIR=1 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "pair or edgeconv_staged" --timeout 200 -p no:cacheprovider 2>&1 | tail -2
for i in 1 2; do
echo "== single"; python scripts/kbench.py k2 2>&1 | grep -v Warning
echo "== pair"; CP_EDGECONV_PAIR=1 python scripts/kbench.py k2 2>&1 | grep -v Warning
done
