# A/B: nanosleep back-off in the idle-role barrier waits (CP_IDLE_SLEEP_NS), whole step under the power cap
for ns in 0 200 500 1000; do
  python - <<PY
from checkerpose_b200.build import build_library
build_library(force=True, defines=["CP_IDLE_SLEEP_NS=$ns"])
PY
  for rep in 1 2; do
  python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('sleep_ns=$ns', 'ms/step', round(d['ms_per_step'],3), 'K2', round(d['roofline']['avg_launch_ms'],4), 'K3', round(d['roofline_k3']['avg_launch_ms'],4), d['clocks']['sm_mhz'])
"
  done
done
