# K2 iteration loop: staged-EdgeConv parity tests, then the micro-benchmark, for the in-tree build and every variant
run() {
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "edgeconv_staged or chain_agg" --timeout 200 -p no:cacheprovider 2>&1 | tail -2
  python scripts/kbench.py k2 2>&1 | grep -v Warning
}
run
for f in variants/lib_*.so; do
  [ -f "$f" ] && CHECKERPOSE_B200_LIB=$PWD/$f run
done
true
