mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu_launch rc=$?"
python scripts/launch_summary.py gpurun_out/launches.csv 3 > gpurun_out/launches.txt 2>&1; head -14 gpurun_out/launches.txt
