mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chain_kernel -s 10 -c 4 -o gpurun_out/prof_chain python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_chain.log 2>&1; echo "ncu_chain rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
