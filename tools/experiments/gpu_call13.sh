#!/bin/bash
# GPU call: host-path time vs archive size (fixed overhead vs per-byte cost).
tag=${1:-r01r}
mkdir -p gpurun_out
for mib in 16 64 256 512 1024 2048; do
  timeout 300 python tools/time_e2e.py $mib 65536 8 size$mib >> gpurun_out/${tag}_e2e.jsonl 2>> gpurun_out/${tag}_e2e.err
done
cat gpurun_out/${tag}_e2e.jsonl
