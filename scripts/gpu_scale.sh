# weak-scaling bench at N GPUs of one box (N = $1)
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench $N rc=$?"
tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-400; tail -3 gpurun_out/bench_${N}gpu.err
