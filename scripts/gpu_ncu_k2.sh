mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:edgeconv_kernel -s 3 -c 1 -f -o gpurun_out/prof_k2 python scripts/kbench.py k2 > gpurun_out/ncu_k2.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_k2.log
python scripts/kbench.py k2 2>&1 | grep -v Warning
