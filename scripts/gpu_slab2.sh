mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu --timeout 300 -p no:cacheprovider -k "slab or border" ) 2>&1 | grep -E "^E  |passed|failed|^FAILED" | head -20
for i in 1 2; do for f in 1 0; do
CP_FUSE_UPSAMPLE=$f python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --image-branch tcgen05 > gpurun_out/ab_f$f.log 2>/dev/null
python - gpurun_out/ab_f$f.log fuse=$f <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "gnn_only", round(d["gnn_only"]["ms_per_step"], 3), "conv", round(d["roofline_conv"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
done; done
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --image-branch cudnn 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cudnn ms/step', round(d['ms_per_step'],3))"
