# failing tests re-run + launch lists (bf16 default, fp32 mode)
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -s ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_fp32.csv python bench.py --dtype fp32 --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch_fp32.log 2>&1; echo "ncu_launch fp32 rc=$?" >> gpurun_out/rc.txt
python scripts/launch_summary.py gpurun_out/launches_fp32.csv 3 > gpurun_out/r02_c_launches_fp32.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu_launch rc=$?" >> gpurun_out/rc.txt
python scripts/launch_summary.py gpurun_out/launches.csv 3 > gpurun_out/r02_c_launches.txt 2>&1
cat gpurun_out/rc.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head; head -30 gpurun_out/r02_c_launches_fp32.txt; head -30 gpurun_out/r02_c_launches.txt
