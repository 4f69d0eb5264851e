mkdir -p gpurun_out
KB_B=64 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_x3 -s 1 -c 1 -o /tmp/prof_x3c -f python scripts/kbench_x3.py conv > gpurun_out/ncu_x3c.log 2>&1; echo "ncu rc=$?"
python scripts/ncu_summary.py /tmp/prof_x3c.ncu-rep 2>&1 | tail -22
python scripts/ncu_hot.py /tmp/prof_x3c.ncu-rep 0 40 2>&1 | cut -c1-190
