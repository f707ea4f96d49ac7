#!/bin/bash
# dev call: decode GPU tests + timing sweep over env settings. usage: tools/gpu_call_c.sh <tag> "<env1>" "<env2>" ...
tag=$1; shift
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_decode.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
for env in "$@"; do
  env $env timeout 300 python tools/time_decode.py 1024 65536 5 "$env" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
done
cat gpurun_out/${tag}_dec.jsonl
