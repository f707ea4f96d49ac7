#!/bin/bash
# round-2 dev call: GPU parity tests, then decode timing sweeps. usage: tools/gpu_call_a.sh <tag>
tag=${1:-r03a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
for env in "" "ZRA_B200_CHUNKS=1" "ZRA_B200_CHUNKS=2" "ZRA_B200_CHUNKS=8" "ZRA_B200_SEQ_CTAS_PER_SM=1" "ZRA_B200_SEQ_CTAS_PER_SM=1 ZRA_B200_CHUNKS=8"; do
  env $env timeout 300 python tools/time_decode.py 1024 65536 5 "$env" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
done
timeout 300 python tools/time_decode.py 1024 16384 5 "16k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 300 python tools/time_decode.py 1024 262144 5 "256k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
cat gpurun_out/${tag}_dec.jsonl
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_decode.py -x -q -k "golden" > gpurun_out/${tag}_sanitizer.log 2>&1; tail -5 gpurun_out/${tag}_sanitizer.log
