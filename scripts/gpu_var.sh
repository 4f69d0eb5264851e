# run scripts/kbench.py against every library variant under variants/
for f in variants/lib_*.so; do
  CHECKERPOSE_B200_LIB=$PWD/$f python scripts/kbench.py $KB_WHAT 2>&1 | grep -v Warning
done
