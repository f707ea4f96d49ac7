#!/bin/bash
tag=${1:-r01h}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
run() { timeout 300 python tools/time_compress.py "$@" >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err; }
for k in mixed text; do for lv in 3 1; do run 256 65536 $lv 3 $k; done; done
cat gpurun_out/${tag}_enc.jsonl; tail -5 gpurun_out/${tag}_enc.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_enc_match_cta -c 1 -o gpurun_out/${tag}_match_l3 \
   python tools/time_compress.py 64 65536 3 1 text > gpurun_out/${tag}_ncu_full.log 2>&1
tail -3 gpurun_out/${tag}_ncu_full.log
