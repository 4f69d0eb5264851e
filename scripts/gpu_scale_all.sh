# scaling on one 8-GPU box: default workload (weak) at 1/2/4/8 and ycbv1024 (strong) at 1/2/4/8
mkdir -p gpurun_out
run() { # name nproc args...
  local name=$1 n=$2; shift 2
  if [ "$n" = "1" ]; then python bench.py --gpus 1 "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; fi
  echo "$name rc=$?"
  python - gpurun_out/$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("  n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "e2e wall ms/step", round(d["e2e"].get("wall_ms_per_step", 0), 2))
except Exception as e:
    print("  unreadable:", e)
PY
}
for n in 1 2 4 8; do run r02_scale_full4096_${n}gpu $n --steps 20 --warmup 3 --no-cpu-baseline --no-parity; done
for n in 1 2 4 8; do run r02_scale_ycbv1024_${n}gpu $n --config ycbv1024 --steps 10 --warmup 3 --no-parity; done
tail -n 2 gpurun_out/r02_scale_*8gpu.err
# BASELINE configs[4] at 8 GPUs (the 1-GPU table is produced by scripts/gpu_hygiene.sh)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --config lm_sweep --steps 5 --warmup 3 > gpurun_out/r02_bench_lm_sweep_8gpu.json 2> gpurun_out/lm_sweep8.err; echo "lm_sweep 8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r02_bench_lm_sweep_8gpu.json").read().strip().splitlines()[-1])
    print("lm_sweep 8 GPUs:", [(t["npoint"], t["graph_k"], round(t["rois_per_s"])) for t in d["sweep"]])
except Exception as e:
    print("unreadable", e)
PY
