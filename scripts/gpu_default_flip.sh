mkdir -p gpurun_out; rm -f gpurun_out/rc.txt
( timeout 1500 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider ) > gpurun_out/t_gpu.log 2>&1; echo "gpu tests rc=$?"
grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -3; grep -E "^FAILED|^E  " gpurun_out/t_gpu.log | head -20
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash scripts/gpu_ab_branch.sh
