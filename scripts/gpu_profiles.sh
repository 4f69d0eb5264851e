# round-2 evidence of the bench command: launch list + ncu --set full of the dominant kernels inside bench.py --profile
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file /tmp/launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu_launch rc=$?"
python scripts/launch_summary.py /tmp/launches.csv 3 > gpurun_out/r02_h_launches.txt 2>&1; head -8 gpurun_out/r02_h_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edgeconv_kernel -s 13 -c 2 -f -o /tmp/prof_ec python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_ec.log 2>&1; echo "ncu ec rc=$?"
python scripts/ncu_summary.py /tmp/prof_ec.ncu-rep > gpurun_out/r02_h_edgeconv.txt 2>&1
python scripts/ncu_hot.py /tmp/prof_ec.ncu-rep 0 30 > gpurun_out/r02_h_edgeconv_hot_instructions.txt 2>&1
timeout 900 ncu --set full --clock-control none -k regex:"taps_chain_kernel|query_tail_kernel" -s 4 -c 2 -f -o /tmp/prof_k3 python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_k3.log 2>&1; echo "ncu k3 rc=$?"
python scripts/ncu_summary.py /tmp/prof_k3.ncu-rep > gpurun_out/r02_h_taps_chain_query_tail.txt 2>&1
grep -E "^## launch|^duration|dram read|dram write|tensor pipe|shared-memory wavefronts %" gpurun_out/r02_h_edgeconv.txt gpurun_out/r02_h_taps_chain_query_tail.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_h_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_h_bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
python bench.py --dtype fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02_h_bench_fp32.json 2>> gpurun_out/bench.err; echo "fp32 rc=$?"
python bench.py --config init64 --steps 20 > gpurun_out/r02_h_bench_init64.json 2>> gpurun_out/bench.err; echo "init64 rc=$?"
tail -n 3 gpurun_out/bench.err; cut -c1-300 gpurun_out/r02_h_bench_reference.json
