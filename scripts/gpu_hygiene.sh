# measurement hygiene: K1 sweep, lm_sweep, ncu summaries of the small kernels (reports summarised on the box and deleted)
mkdir -p gpurun_out
python scripts/kbench_knn.py 2>&1 | grep -v Warning > gpurun_out/r02_knn_sweep.txt; tail -3 gpurun_out/r02_knn_sweep.txt
( timeout 1200 python bench.py --config lm_sweep --steps 5 --warmup 3 ) > gpurun_out/r02_bench_lm_sweep_1gpu.json 2> gpurun_out/lm_sweep.err; echo "lm_sweep rc=$?"; tail -n 2 gpurun_out/lm_sweep.err
timeout 900 ncu --set full --clock-control none -k regex:"knn_kernel|decode_|correspondences|chain_kernel|upsample2x|transpose_|bias_add|permute_rows" -c 30 -o /tmp/prof_small -f python bench.py --profile --steps 1 --warmup 1 > gpurun_out/ncu_small.log 2>&1; echo "ncu small rc=$?"
python scripts/ncu_summary.py /tmp/prof_small.ncu-rep > gpurun_out/r02_small_kernels.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"sample_taps|edge_aggregate_staged|gemm_x3" -c 14 -o /tmp/prof_fp32 -f python bench.py --dtype fp32 --batch 64 --profile --steps 1 --warmup 1 > gpurun_out/ncu_fp32.log 2>&1; echo "ncu fp32 rc=$?"
python scripts/ncu_summary.py /tmp/prof_fp32.ncu-rep > gpurun_out/r02_fp32_kernels.txt 2>&1
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "edgeconv_staged or taps_chain_fused or x3 or edge_aggregate_staged or correspondences_packed" ) > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -n 4 gpurun_out/r02_sanitizer_memcheck.txt
( timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -q -m gpu -p no:cacheprovider -k "edgeconv_staged or taps_chain_fused or x3_linear or edge_aggregate_staged" ) > gpurun_out/r02_sanitizer_racecheck.txt 2>&1; echo "racecheck rc=$?"; grep -E "Race reported|RACECHECK SUMMARY" gpurun_out/r02_sanitizer_racecheck.txt | sort | uniq -c | sort -rn | head -40
ls -la gpurun_out | head -30
