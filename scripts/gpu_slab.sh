mkdir -p gpurun_out
for pair in 0 1; do
  KB_CHECK_ONLY=1 CP_SLAB_PAIR=$pair timeout 120 python scripts/kbench_slab.py 2>&1 | grep -v Warning
done | tee gpurun_out/slab_check.txt
timeout 300 python scripts/kbench_slab.py 2>&1 | grep -E "^conv" | tee gpurun_out/kbench_slab.txt
echo "== single CTA"; CP_SLAB_PAIR=0 timeout 300 python scripts/kbench_slab.py 2>&1 | grep -E "^conv" | tee gpurun_out/kbench_slab_single.txt
