#!/bin/bash
# GPU call: parity tests, full bench (with compress leg), chunk-count sweep.
tag=${1:-r01d}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
tail -3 gpurun_out/${tag}_bench.err
for c in 1 2 3 4 6 8 12 16 32; do
  ZRA_B200_CHUNKS=$c timeout 120 python tools/time_decode.py 1024 65536 5 chunks$c >> gpurun_out/${tag}_sweep.jsonl 2>> gpurun_out/${tag}_sweep.err
done
cat gpurun_out/${tag}_sweep.jsonl
