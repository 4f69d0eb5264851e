mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -s ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head -20
