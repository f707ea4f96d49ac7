#!/bin/bash
# GPU call: producer/consumer matcher (ZRA_B200_ENC_PIPE=1), guarded by short timeouts (a barrier bug would hang).
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 60 python tools/time_compress.py 16 65536 3 1 text > gpurun_out/${tag}_first.json 2> gpurun_out/${tag}_first.err; echo "first rc=$?"; cat gpurun_out/${tag}_first.json | cut -c1-200; tail -2 gpurun_out/${tag}_first.err
if grep -q '"ok": true' gpurun_out/${tag}_first.json; then
for a in "65536 3 text" "65536 3 mixed" "65536 1 text" "65536 1 mixed" "16384 3 text" "65536 2 text"; do set -- $a
  timeout 120 python tools/time_compress.py 256 $1 $2 3 $3 >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err
  timeout 120 python tools/time_compress.py 256 $1 $2 3 $3 >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err
done
cut -c1-230 gpurun_out/${tag}_enc.jsonl
timeout 300 python -m pytest tests/test_gpu_encode.py -m gpu -q -x 2>&1 | tail -3
fi
