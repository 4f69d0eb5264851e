mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head -20
for ib in tcgen05 cudnn; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --image-branch $ib > gpurun_out/bench_$ib.log 2> gpurun_out/bench.err; echo "bench $ib rc=$?"
python - gpurun_out/bench_$ib.log <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
      "K2 ms", round(d["roofline"]["avg_launch_ms"], 4), "frac", round(d["roofline"]["frac"], 3),
      "K3 ms", round(d["roofline_k3"]["avg_launch_ms"], 4), "gnn_only", round(d["gnn_only"]["ms_per_step"],3), "conv", (d.get("roofline_conv") or {}).get("ms_per_step"), (d.get("roofline_conv") or {}).get("achieved"), "parity", d["parity"]["keypoint_agreement"], d["clocks"], "launches", d["gpu_launches"])
PY
done
tail -n 3 gpurun_out/bench.err
