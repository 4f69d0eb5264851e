mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file /tmp/launches.csv python bench.py --profile --steps 2 --warmup 1 $BENCH_ARGS > gpurun_out/ncu_launch.log 2>&1; echo "ncu_launch rc=$?"
python scripts/launch_summary.py /tmp/launches.csv 3 > gpurun_out/launches.txt 2>&1; head -40 gpurun_out/launches.txt
python - <<'PY'
import csv
lines=[l for l in open('/tmp/launches.csv') if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if 'conv_' in r['Kernel Name'] or 'upsample' in r['Kernel Name'] or 'zero_border' in r['Kernel Name']]
per=rows[-22:]
print([(r['Kernel Name'][11:20], round(float(r['Metric Value'].replace(',',''))/1e6,3)) for r in per])
PY
