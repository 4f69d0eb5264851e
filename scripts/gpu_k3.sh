run() {
  timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "taps_chain or chain" --timeout 200 -p no:cacheprovider 2>&1 | tail -2
  python scripts/kbench.py k2 k3 2>&1 | grep -v Warning | grep K3
}
run
for f in variants/lib_*.so; do
  [ -f "$f" ] && CHECKERPOSE_B200_LIB=$PWD/$f run
done
true
