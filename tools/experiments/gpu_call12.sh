#!/bin/bash
# GPU call: host path A/B within one box: ordered copy streams vs copies on the chunk streams, several chunk counts, repeated.
tag=${1:-r01p}
mkdir -p gpurun_out
python tools/pcie_probe.py 1024 > gpurun_out/${tag}_pcie.json; cat gpurun_out/${tag}_pcie.json
e2e() { name=$1; shift; env "$@" timeout 200 python tools/time_e2e.py 1024 65536 8 $name >> gpurun_out/${tag}_e2e.jsonl 2>> gpurun_out/${tag}_e2e.err; }
for rep in 1 2; do
e2e ord4 A=1
e2e un4 ZRA_B200_UNORDERED_IO=1
e2e ord8 ZRA_B200_IO_CHUNKS=8
e2e un8 ZRA_B200_UNORDERED_IO=1 ZRA_B200_IO_CHUNKS=8
e2e ord16 ZRA_B200_IO_CHUNKS=16
e2e un16 ZRA_B200_UNORDERED_IO=1 ZRA_B200_IO_CHUNKS=16
done
cat gpurun_out/${tag}_e2e.jsonl

