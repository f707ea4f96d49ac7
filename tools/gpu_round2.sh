#!/bin/bash
# One gpurun call for the record: GPU parity tests, bench line (both arms), ncu launch list, ncu full capture of the decode kernels.
# usage: tools/gpu_round2.sh <tag>   (outputs under gpurun_out/<tag>_*; summarise with tools/summarise_profiles.py <tag>)
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -3 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2>> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${tag}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ra --no-compress --no-streaming > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ZRA_B200_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(block_setup|huf_decode|seq_decode|seq_redo|seq_execute|frame_finish)' -c 6 \
    -f -o gpurun_out/${tag}_full python tools/profile_decode.py 1024 65536 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${tag}_full.ncu-rep
