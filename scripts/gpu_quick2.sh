mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 200 -p no:cacheprovider -k "upsample" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED|Error" | head
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:upsample -c 6 --csv --log-file /tmp/l.csv python bench.py --profile --steps 1 --warmup 1 > /dev/null 2>&1; grep upsample /tmp/l.csv | awk -F'","' '{print $NF}' | tr -d '"' | tail -4
