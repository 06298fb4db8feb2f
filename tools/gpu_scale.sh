#!/bin/bash
# weak-scaling check at N GPUs (the driver's own launch line), N from $1
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.log 2>&1; echo "bench n=$N exit $?"
tail -2 gpurun_out/bench_n$N.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref n=$N exit $?"
tail -1 gpurun_out/bench_ref_n$N.log | cut -c1-200
