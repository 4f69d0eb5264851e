mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -k "x3" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head
python scripts/kbench_x3.py 2>&1 | grep -v Warning | tee gpurun_out/kbench_x3.txt
( timeout 600 python bench.py --dtype fp32 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_fp32.json").read().strip().splitlines()[-1])
print(" ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "gnn_only", (d.get("gnn_only") or {}).get("ms_per_step"), "parity", d["parity"]["keypoint_agreement"])
PY
