# round-2 (k) evidence with the slab image branch as default: launch list + ncu --set full of the slab convolutions + bench lines
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file /tmp/launches.csv python bench.py --profile --steps 2 --warmup 1 > gpurun_out/ncu_launch.log 2>&1; echo "ncu_launch rc=$?"
python scripts/launch_summary.py /tmp/launches.csv 3 > gpurun_out/r02_k_launches.txt 2>&1; head -12 gpurun_out/r02_k_launches.txt
# the stage-2 convolutions (512->256 and 256->256 at 64x64) are the 8th and 9th conv_slab launches of a step (after the parities, 2+1, 2+1)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_slab_kernel -s 18 -c 4 -f -o /tmp/prof_cs python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_cs.log 2>&1; echo "ncu cs rc=$?"
python scripts/ncu_summary.py /tmp/prof_cs.ncu-rep > gpurun_out/r02_k_conv_slab.txt 2>&1
python scripts/ncu_hot.py /tmp/prof_cs.ncu-rep 0 20 > gpurun_out/r02_k_conv_slab_hot_instructions.txt 2>&1
grep -E "^## launch|^duration|dram read|dram write|tensor pipe|shared-memory wavefronts %|issue slots" gpurun_out/r02_k_conv_slab.txt
cuobjdump -sass checkerpose_b200/csrc/libcheckerpose_b200.so 2>/dev/null | grep -E "UTMALDG|UTCHMMA|UTMASTG|UTCBAR" | sed 's/^ *//' | cut -c1-60 | sort | uniq -c | sort -rn | head -12 > gpurun_out/r02_k_sass_mnemonics.txt; cat gpurun_out/r02_k_sass_mnemonics.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_k_bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --image-branch cudnn > gpurun_out/r02_k_bench_cudnn_arm.json 2>> gpurun_out/bench.err; echo "bench cudnn rc=$?"
python bench.py --config init64 --steps 20 --no-cpu-baseline > gpurun_out/r02_k_bench_init64.json 2>> gpurun_out/bench.err; echo "init64 rc=$?"
python bench.py --config ycbv1024 --steps 10 --no-cpu-baseline > gpurun_out/r02_k_bench_ycbv1024.json 2>> gpurun_out/bench.err; echo "ycbv rc=$?"
tail -n 3 gpurun_out/bench.err
python - <<'PY'
import json
for f in ("r02_k_bench", "r02_k_bench_cudnn_arm", "r02_k_bench_init64", "r02_k_bench_ycbv1024"):
    d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    print(f, round(d["ms_per_step"], 3), round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 3) if d.get("roofline") else None, d["clocks"])
PY
