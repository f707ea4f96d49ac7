#!/bin/bash
# dev call: GPU parity tests (decode only, fast) + decode timing. usage: tools/gpu_call_b.sh <tag> [full]
tag=${1:-r03c}
mkdir -p gpurun_out
if [ "$2" = "full" ]; then
  timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
else
  timeout 1200 python -m pytest tests/test_gpu_decode.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
fi
tail -4 gpurun_out/${tag}_pytest.log
for env in "" "ZRA_B200_CHUNKS=1" "ZRA_B200_CHUNKS=2" "ZRA_B200_CHUNKS=8"; do
  env $env timeout 300 python tools/time_decode.py 1024 65536 5 "$env" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
done
timeout 300 python tools/time_decode.py 1024 16384 5 "16k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 300 python tools/time_decode.py 1024 262144 5 "256k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
cat gpurun_out/${tag}_dec.jsonl
