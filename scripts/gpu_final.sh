# final regression of the round: whole GPU suite, smoke, every bench mode once
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/bench.err; echo "ref rc=$?"
python bench.py --dtype fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_fp32.json 2>> gpurun_out/bench.err; echo "fp32 rc=$?"
python bench.py --config lm_sweep --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_lm_sweep.json 2>> gpurun_out/bench.err; echo "lm_sweep rc=$?"
python bench.py --config init64 --steps 20 --no-cpu-baseline > gpurun_out/final_bench_init64.json 2>> gpurun_out/bench.err; echo "init64 rc=$?"
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("final_bench", "final_bench_reference", "final_bench_fp32", "final_bench_init64"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d.items() if k in ("impl", "value", "ms_per_step", "gpu_launches", "dtype")}, "e2e", round(d["e2e"]["value"], 1), "parity", (d.get("parity") or {}).get("keypoint_agreement"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
d = json.loads(open("gpurun_out/final_bench_lm_sweep.json").read().strip().splitlines()[-1])
print("lm_sweep:", [(t["npoint"], t["graph_k"], round(t["rois_per_s"])) for t in d["sweep"]])
PY
