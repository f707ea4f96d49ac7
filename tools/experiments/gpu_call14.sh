#!/bin/bash
# GPU call: sequence-stage slot counts at large archive sizes (does co-residency of the execute kernel pay when throughput, not latency, is the bound?)
tag=${1:-r01v}
mkdir -p gpurun_out
run() { name=$1; mib=$2; shift 2; env "$@" timeout 400 python tools/time_decode.py $mib 65536 3 $name >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err; }
for mib in 4096 1024; do
run s88_$mib $mib A=1
run s72_$mib $mib ZRA_B200_SEQ_SLOTS=72
run s58_$mib $mib ZRA_B200_SEQ_SLOTS=58
run s44_$mib $mib ZRA_B200_SEQ_SLOTS=44
run s58c8_$mib $mib ZRA_B200_SEQ_SLOTS=58 ZRA_B200_CHUNKS=8
run s88c8_$mib $mib ZRA_B200_CHUNKS=8
run s88c16_$mib $mib ZRA_B200_CHUNKS=16
done
cut -c1-120 gpurun_out/${tag}_dec.jsonl
