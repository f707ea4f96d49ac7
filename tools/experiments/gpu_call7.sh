#!/bin/bash
# GPU call: exec-kernel prefetch / occupancy variants and chunk plans (device-resident decode of the bench archive).
tag=${1:-r01f}
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python tools/time_decode.py 1024 65536 5 $name >> gpurun_out/${tag}_dec.jsonl 2>> gpurun_out/${tag}_dec.err; }
run base A=1
run pf1 ZRA_B200_EXEC_FLAGS=1
run pf2 ZRA_B200_EXEC_FLAGS=2
run pf3 ZRA_B200_EXEC_FLAGS=3
run occ5 ZRA_B200_EXEC_OCC=5
run occ5pf3 ZRA_B200_EXEC_OCC=5 ZRA_B200_EXEC_FLAGS=3
run planA ZRA_B200_PLAN=13024
run planB ZRA_B200_PLAN=3360,13024
run planC ZRA_B200_PLAN=6512,6512
run planD ZRA_B200_PLAN=4342,4341,4341
run planApf3 ZRA_B200_PLAN=13024 ZRA_B200_EXEC_FLAGS=3
cat gpurun_out/${tag}_dec.jsonl | cut -c1-400
ZRA_B200_TIMELINE=1 ZRA_B200_PLAN=13024 timeout 200 python tools/timeline.py 1024 65536 2> gpurun_out/${tag}_timeline_planA.txt
