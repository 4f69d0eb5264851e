# in-step A/B of K2's CTA-pair variant (CP_EDGECONV_PAIR=1) against the single-CTA kernel, alternating
mkdir -p gpurun_out
for i in 1 2; do for pr in 0 1; do
CP_EDGECONV_PAIR=$pr python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/ab_pair$pr.log 2>/dev/null
python - gpurun_out/ab_pair$pr.log pair=$pr <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "K2", round(d["roofline"]["avg_launch_ms"], 4), "gnn_only", round(d["gnn_only"]["ms_per_step"], 3), "clk", d["clocks"]["sm_mhz"])
PY
done; done | tee gpurun_out/ab_pair.txt
