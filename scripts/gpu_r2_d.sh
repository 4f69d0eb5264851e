mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -s ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
( timeout 600 python bench.py --dtype fp32 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_d_bench_fp32.json 2> gpurun_out/bench_fp32b.err; echo "bench fp32 256 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_d_bench_fp32.json").read().strip().splitlines()[-1])
print(" ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gnn_only", (d.get("gnn_only") or {}).get("ms_per_step"), "parity", d.get("parity"))
PY
tail -n 3 gpurun_out/bench_fp32b.err
