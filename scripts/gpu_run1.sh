mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python __graft_entry__.py > gpurun_out/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "not chain and not bf16" --timeout 300 -p no:cacheprovider > gpurun_out/t_kernels.log 2>&1; echo "kernels rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "chain_gemm_exact" --timeout 120 -p no:cacheprovider > gpurun_out/t_chain_gemm.log 2>&1; echo "chain_gemm rc=$?" >> gpurun_out/rc.txt
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "chain_three or chain_agg or bf16" --timeout 120 -p no:cacheprovider > gpurun_out/t_chain_other.log 2>&1; echo "chain_other rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_head.py -q -m gpu -k "fp32 or lm_per" --timeout 300 -p no:cacheprovider > gpurun_out/t_head_fp32.log 2>&1; echo "head_fp32 rc=$?" >> gpurun_out/rc.txt
timeout 900 python -m pytest tests/test_gpu_head.py -q -m gpu -k "bf16 or full_size" -s --timeout 300 -p no:cacheprovider > gpurun_out/t_head_bf16.log 2>&1; echo "head_bf16 rc=$?" >> gpurun_out/rc.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rc.txt
cat gpurun_out/rc.txt
