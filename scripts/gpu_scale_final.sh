# final scaling run (slab image branch as default): default workload (weak) at 1/2/4/8, ycbv1024 (strong) at 1 and 8
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
    print("  n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 3), "RoIs/s", round(d["value"]), "e2e", round(d["e2e"]["value"]), "launches", d["gpu_launches"])
except Exception as e:
    print("  unreadable:", e)
PY
}
for n in 8 4 2 1; do run r02_k_scale_full4096_${n}gpu $n --steps 20 --warmup 3 --no-cpu-baseline --no-parity; done
for n in 8 1; do run r02_k_scale_ycbv1024_${n}gpu $n --config ycbv1024 --steps 10 --warmup 3 --no-cpu-baseline --no-parity; done
tail -n 2 gpurun_out/r02_k_scale_*8gpu.err
