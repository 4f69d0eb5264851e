mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taps_chain_kernel -s 2 -c 1 -f -o /tmp/prof_k3 python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
python scripts/ncu_hot.py /tmp/prof_k3.ncu-rep 0 60 > gpurun_out/r02_k_taps_chain_hot_instructions.txt 2>&1
head -80 gpurun_out/r02_k_taps_chain_hot_instructions.txt
