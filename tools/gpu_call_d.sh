#!/bin/bash
# full GPU test suite + decode timings at several shapes. usage: tools/gpu_call_d.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 300 python tools/time_decode.py 1024 65536 5 "64k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 300 python tools/time_decode.py 1024 16384 5 "16k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 300 python tools/time_decode.py 1024 262144 5 "256k" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 600 python tools/time_decode.py 4096 65536 3 "4g" >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
cut -c1-330 gpurun_out/${tag}_dec.jsonl
