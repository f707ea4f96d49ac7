#!/bin/bash
# GPU call: host-path chunk plans (e2e) and encoder table defaults (levels 2/3, 16/64 KiB frames) + encoder tests.
tag=${1:-r01h}
mkdir -p gpurun_out
e2e() { name=$1; shift; env "$@" timeout 200 python tools/time_e2e.py 1024 65536 5 $name >> gpurun_out/${tag}_e2e.jsonl 2>> gpurun_out/${tag}_e2e.err; }
e2e geo A=1
e2e io4 ZRA_B200_IO_CHUNKS=4
e2e p1 ZRA_B200_IO_PLAN=256,512,1024,2048,4096
e2e p2 ZRA_B200_IO_PLAN=1024,2048,4096,4096
e2e p3 ZRA_B200_IO_PLAN=512,1024,2048,2048,2048,2048,2048,2048
cat gpurun_out/${tag}_e2e.jsonl
enc() { fs=$1; lv=$2; kind=$3; shift 3; env "$@" timeout 300 python tools/time_compress.py 256 $fs $lv 3 $kind >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err; }
enc 65536 3 text A=1
enc 65536 3 mixed A=1
enc 65536 2 text A=1
enc 65536 2 text ZRA_B200_ENC_LOGS=13
enc 65536 2 text ZRA_B200_ENC_LOGS=14
enc 65536 1 text A=1
enc 16384 3 text A=1
enc 16384 3 text ZRA_B200_ENC_LOGS=14 ZRA_B200_ENC_LOGL=15
enc 16384 3 text ZRA_B200_ENC_LOGS=11 ZRA_B200_ENC_LOGL=12
enc 16384 1 text A=1
cut -c1-330 gpurun_out/${tag}_enc.jsonl; tail -3 gpurun_out/${tag}_enc.err
timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
