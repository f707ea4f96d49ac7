#!/bin/bash
# GPU call: shared-memory matcher table-size / thread sweep for level 3 (speed and archive size).
tag=${1:-r01g}
mkdir -p gpurun_out
run() { kind=$1; shift; env "$@" timeout 300 python tools/time_compress.py 256 65536 3 3 $kind >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err; }
for kind in text mixed; do
run $kind A=1
run $kind ZRA_B200_ENC_LOGS=14 ZRA_B200_ENC_LOGL=15
run $kind ZRA_B200_ENC_LOGS=14 ZRA_B200_ENC_LOGL=15 ZRA_B200_ENC_THREADS=256
run $kind ZRA_B200_ENC_LOGS=13 ZRA_B200_ENC_LOGL=15 ZRA_B200_ENC_THREADS=256
run $kind ZRA_B200_ENC_LOGS=14 ZRA_B200_ENC_LOGL=14 ZRA_B200_ENC_THREADS=256
run $kind ZRA_B200_ENC_LOGS=13 ZRA_B200_ENC_LOGL=14 ZRA_B200_ENC_THREADS=256
run $kind ZRA_B200_ENC_LOGS=13 ZRA_B200_ENC_LOGL=13 ZRA_B200_ENC_THREADS=256
done
cut -c1-330 gpurun_out/${tag}_enc.jsonl; tail -3 gpurun_out/${tag}_enc.err
