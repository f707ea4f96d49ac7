#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list, ncu full capture of the decode kernels.
# usage: tools/gpu_round.sh <tag>   (outputs under gpurun_out/<tag>_*)
tag=${1:-rXX}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -5 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cat gpurun_out/${tag}_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2>> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ra > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
ZRA_B200_CHUNKS=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(block_setup|huf_decode|seq_decode|seq_execute|frame_finish)' -c 5 \
    -f -o gpurun_out/${tag}_full python tools/profile_decode.py 1024 65536 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
# encoder and random-access kernels: full-set capture, exported to raw CSV on the box (the reports are too big to bring back)
timeout 600 ncu --set full --clock-control none -k regex:'k_enc_|k_scan_|k_write_entries|k_gather_frames|k_crc_' -c 16 \
    -f -o /tmp/${tag}_enc_full python tools/time_compress.py 1024 65536 3 1 mixed > gpurun_out/${tag}_ncu_enc_full.log 2>&1; echo "ncu enc full rc=$?"
ncu -i /tmp/${tag}_enc_full.ncu-rep --page raw --csv > gpurun_out/${tag}_enc_full_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k regex:'k_ra_|k_block_setup|k_huf_decode|k_seq_decode|k_seq_execute|k_frame_finish' -c 12 \
    -f -o /tmp/${tag}_ra_full python tools/profile_ra.py 1024 1 > gpurun_out/${tag}_ncu_ra_full.log 2>&1; echo "ncu ra full rc=$?"
ncu -i /tmp/${tag}_ra_full.ncu-rep --page raw --csv > gpurun_out/${tag}_ra_full_raw.csv 2>/dev/null
du -sh gpurun_out
ls -la gpurun_out
