#!/bin/bash
# GPU call: forked Huffman stage / slot counts / stream priorities (device-resident + host path) + decode tests.
tag=${1:-r01i}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python tools/time_decode.py 1024 65536 5 $name >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err; }
run base A=1
run flat ZRA_B200_FLAT_PRIORITY=1
run fork78 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=78
run fork72 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=72
run fork64 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=64
run fork64c2 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=64 ZRA_B200_CHUNKS=2
run fork64c8 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=64 ZRA_B200_CHUNKS=8
run fork88 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=88
run s64 ZRA_B200_SEQ_SLOTS=64
cut -c1-200 gpurun_out/${tag}_dec.jsonl
ZRA_B200_TIMELINE=1 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=64 timeout 200 python tools/timeline.py 1024 65536 2> gpurun_out/${tag}_timeline_fork64.txt
e2e() { name=$1; shift; env "$@" timeout 200 python tools/time_e2e.py 1024 65536 5 $name >> gpurun_out/${tag}_e2e.jsonl 2>> gpurun_out/${tag}_e2e.err; }
e2e prio4 A=1
e2e flat4 ZRA_B200_FLAT_PRIORITY=1
e2e geo ZRA_B200_IO_GEOMETRIC=1
e2e io8 ZRA_B200_IO_CHUNKS=8
e2e fork64 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=64
e2e fork64io8 ZRA_B200_FORK_HUF=1 ZRA_B200_SEQ_SLOTS=64 ZRA_B200_IO_CHUNKS=8
cat gpurun_out/${tag}_e2e.jsonl
ZRA_B200_FORK_HUF=1 timeout 900 python -m pytest tests/test_gpu_decode.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
