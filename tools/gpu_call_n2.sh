#!/bin/bash
# GPU call (--gpus 2): the N=2 bench exactly as the driver launches it (NCCL: frame-size exchange, all-gather leg).
tag=${1:-r01n2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.txt 2>&1
NCCL_DEBUG=WARN timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json | cut -c1-3000; tail -5 gpurun_out/${tag}_bench.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
cat gpurun_out/${tag}_bench_ref.json | cut -c1-600
