# A/B of the two image branches on one box, alternating (the board's power cap makes single runs drift)
mkdir -p gpurun_out
for i in 1 2 3; do for ib in cudnn tcgen05; do
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --image-branch $ib > gpurun_out/ab_$ib.log 2>/dev/null
python - gpurun_out/ab_$ib.log $ib <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "gnn_only", round(d["gnn_only"]["ms_per_step"], 3), "K2", round(d["roofline"]["avg_launch_ms"], 4), "clk", d["clocks"]["sm_mhz"])
PY
done; done | tee gpurun_out/ab_branch.txt
