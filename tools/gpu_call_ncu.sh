#!/bin/bash
# ncu --set full capture of the decode kernels (one launch each, whole archive as one chunk). usage: tools/gpu_call_ncu.sh <tag> [frame]
tag=${1:-r03n}
frame=${2:-65536}
mkdir -p gpurun_out
ZRA_B200_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(block_setup|huf_decode|seq_decode|seq_redo|seq_execute|frame_finish)' -c 7 \
    -f -o gpurun_out/${tag}_full python tools/profile_decode.py 1024 $frame 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${tag}_ncu_full.log
ls -la gpurun_out/${tag}_full.ncu-rep
