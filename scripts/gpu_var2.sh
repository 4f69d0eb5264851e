for f in default variants/lib_g2.so; do
  if [ "$f" = "default" ]; then timeout 300 python scripts/kbench.py k2 2>&1 | grep -v Warning | grep K2
  else CHECKERPOSE_B200_LIB=$PWD/$f timeout 300 python scripts/kbench.py k2 2>&1 | grep -v Warning | grep K2; CHECKERPOSE_B200_LIB=$PWD/$f timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -p no:cacheprovider -k "edgeconv_staged" 2>&1 | tail -1; fi
done
