mkdir -p gpurun_out
python scripts/kbench_pnp.py 256 4096 2>&1 | grep -v Warning | tee gpurun_out/r02_pnp_bench.txt
python scripts/kbench_pnp.py 256 512 2>&1 | grep -v Warning | tee -a gpurun_out/r02_pnp_bench.txt
