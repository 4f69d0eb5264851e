# x2 upsampling micro-benchmark for the in-tree build and every variants/lib_*.so
python scripts/kbench_up.py 2>&1 | grep -v Warning
for f in variants/lib_*.so; do [ -f "$f" ] && { echo $f; CHECKERPOSE_B200_LIB=$PWD/$f python scripts/kbench_up.py 2>&1 | grep -v Warning; }; done
true
