mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file /tmp/launches32.csv python bench.py --dtype fp32 --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch32.log 2>&1; echo "ncu_launch rc=$?"
python scripts/launch_summary.py /tmp/launches32.csv 3 > gpurun_out/r02_h_launches_fp32.txt 2>&1; head -12 gpurun_out/r02_h_launches_fp32.txt
python - <<'PY'
import csv
lines=[l for l in open('/tmp/launches32.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if 'gemm_x3' in r['Kernel Name'] or 'edge_aggregate' in r['Kernel Name']]
per=rows[-50:]
print([round(float(r['Metric Value'].replace(',',''))/1e6,3) for r in per])
PY
