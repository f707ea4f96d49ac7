#!/bin/bash
# GPU call: decode parity tests + device-resident timing at the three frame sizes.
tag=${1:-r01j}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q -x > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
run() { name=$1; fs=$2; shift 2; env "$@" timeout 200 python tools/time_decode.py 1024 $fs 5 $name >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err; }
run base 65536 A=1
run f16k 16384 A=1
run f256k 262144 A=1
cut -c1-400 gpurun_out/${tag}_dec.jsonl
ZRA_B200_TIMELINE=1 timeout 200 python tools/timeline.py 1024 65536 2> gpurun_out/${tag}_timeline.txt
