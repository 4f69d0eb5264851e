# round-1 evidence run: smoke, GPU tests, both bench arms, ncu launch list, full captures of K2 and K3
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu_launch rc=$?" >> gpurun_out/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:edgeconv_kernel -s 14 -c 2 -f -o gpurun_out/prof_ec python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?" >> gpurun_out/rc.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:taps_chain_kernel -s 4 -c 2 -f -o gpurun_out/prof_taps python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_taps.log 2>&1; echo "ncu_taps rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -3 gpurun_out/smoke.log; tail -3 gpurun_out/t_gpu.log; cat gpurun_out/bench.log | cut -c1-300; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_ref.log | cut -c1-300
