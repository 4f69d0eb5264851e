mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
python __graft_entry__.py > gpurun_out/build.log 2>&1
( time timeout 900 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider --durations=8 ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
( time timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; tail -15 gpurun_out/t_gpu.log; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
