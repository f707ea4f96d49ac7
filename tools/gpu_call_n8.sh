#!/bin/bash
# GPU call (--gpus 8): the N=8 bench exactly as the driver launches it.
tag=${1:-r01n8}
n=${2:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.txt 2>&1
free -g | head -2 > gpurun_out/${tag}_mem.txt; nproc >> gpurun_out/${tag}_mem.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_mem.txt
