# round 2, first GPU call: the whole GPU test suite, the default bench line, and first lines of the other configs
mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
python -c "import __graft_entry__ as g; g.build(force=False)" > gpurun_out/build.log 2>&1
( time timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider --durations=12 -s ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?" >> gpurun_out/rc.txt
( time timeout 600 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_a_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/rc.txt
( timeout 600 python bench.py --dtype fp32 --batch 64 --steps 3 --warmup 3 --no-cpu-baseline ) > gpurun_out/r02_a_bench_fp32_simt.json 2> gpurun_out/bench_fp32.err; echo "bench fp32 rc=$?" >> gpurun_out/rc.txt
( timeout 600 python bench.py --config init64 --steps 20 ) > gpurun_out/r02_a_bench_init64.json 2> gpurun_out/bench_init64.err; echo "bench init64 rc=$?" >> gpurun_out/rc.txt
( timeout 900 python bench.py --config ycbv1024 --steps 5 ) > gpurun_out/r02_a_bench_ycbv1024.json 2> gpurun_out/bench_ycbv.err; echo "bench ycbv rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt; grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -5
for f in gpurun_out/r02_a_bench*.json; do echo $f; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(" ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "K2 ms", r.get("avg_launch_ms"), "frac", r.get("frac"),
          "gnn_only", (d.get("gnn_only") or {}).get("ms_per_step"), "parity", (d.get("parity") or {}).get("keypoint_agreement"))
except Exception as e:
    print(" unreadable:", e)
PY
done
tail -3 gpurun_out/bench*.err
