#!/bin/bash
# GPU call: PCIe probe, chunk-pipeline timeline, host-path chunk sweep.
tag=${1:-r01e}
mkdir -p gpurun_out
timeout 120 python tools/pcie_probe.py 1024 > gpurun_out/${tag}_pcie.json 2> gpurun_out/${tag}_pcie.err; cat gpurun_out/${tag}_pcie.json
for c in 4 8; do
  ZRA_B200_TIMELINE=1 ZRA_B200_CHUNKS=$c timeout 200 python tools/timeline.py 1024 65536 2> gpurun_out/${tag}_timeline_c$c.txt
done
for c in 4 8 16 32 64; do
  ZRA_B200_IO_CHUNKS=$c timeout 200 python tools/time_e2e.py 1024 65536 5 io$c >> gpurun_out/${tag}_e2e.jsonl 2>> gpurun_out/${tag}_e2e.err
done
cat gpurun_out/${tag}_e2e.jsonl
timeout 200 python tools/time_decode.py 1024 65536 5 base >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 200 python tools/time_decode.py 1024 16384 5 f16k >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
timeout 200 python tools/time_decode.py 1024 262144 5 f256k >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err
cat gpurun_out/${tag}_dec.jsonl
