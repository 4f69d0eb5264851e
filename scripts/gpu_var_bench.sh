# full bench.py for every library variant under variants/ (A/B under the real, power-capped step)
mkdir -p gpurun_out
for rep in 1 2; do
for f in variants/lib_*.so; do
  CHECKERPOSE_B200_LIB=$PWD/$f python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$f', 'ms/step', round(d['ms_per_step'],3), 'K2', round(d['roofline']['avg_launch_ms'],4), 'K3', round(d['roofline_k3']['avg_launch_ms'],4), 'e2e', round(d['e2e']['value']), d['clocks'])"
done
done
