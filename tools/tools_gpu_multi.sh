# bench.py under torchrun on N GPUs (arg 1), plus the 2-rank NCCL path
N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 6 --warmup 3 --no-cpu-baseline --no-pruned-run > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-700; tail -3 gpurun_out/bench_${N}gpu.err
