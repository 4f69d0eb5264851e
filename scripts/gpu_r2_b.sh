# x3 GEMM / conv tests, fp32 head tests, fp32 bench
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -x -s -k "x3" ) > gpurun_out/t_x3.log 2>&1; echo "x3 tests rc=$?" >> gpurun_out/rc.txt
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider --durations=12 -s ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
( timeout 600 python bench.py --dtype fp32 --batch 64 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_b_bench_fp32_b64.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?" >> gpurun_out/rc.txt
( timeout 600 python bench.py --dtype fp32 --steps 5 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_b_bench_fp32.json 2> gpurun_out/bench_fp32b.err; echo "bench fp32 256 rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; grep -E "passed|failed|error" gpurun_out/t_x3.log gpurun_out/t_gpu.log | tail -5; grep -E "^x3 " gpurun_out/t_x3.log
for f in gpurun_out/r02_b_bench*.json; do echo $f; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(" ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]),
          "gnn_only", (d.get("gnn_only") or {}).get("ms_per_step"), "parity", (d.get("parity") or {}).get("keypoint_agreement"))
except Exception as e:
    print(" unreadable:", e)
PY
done
tail -n 3 gpurun_out/bench_fp32.err gpurun_out/bench_fp32b.err
