#!/bin/bash
# GPU call: encoder tests + timing + table-size experiments + launch list
tag=${1:-r01g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
run() { timeout 300 python tools/time_compress.py "$@" >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err; }
for k in mixed text; do for lv in 3 1; do run 256 65536 $lv 3 $k; done; done
run 256 16384 3 3 text
ZRA_B200_ENC_LOGS=14 run 256 65536 1 3 text
ZRA_B200_ENC_LOGS=15 run 256 65536 1 3 text
ZRA_B200_ENC_LOGS=15 ZRA_B200_ENC_MLS=5 run 256 65536 1 3 text
ZRA_B200_ENC_LOGS=13 ZRA_B200_ENC_LOGL=14 run 256 65536 1 3 text
ZRA_B200_ENC_LOGS=14 ZRA_B200_ENC_LOGL=15 run 256 65536 3 3 text
ZRA_B200_ENC_LOGS=13 ZRA_B200_ENC_LOGL=14 run 256 65536 3 3 text
ZRA_B200_ENC_THREADS=256 run 256 65536 3 3 text
cat gpurun_out/${tag}_enc.jsonl; tail -5 gpurun_out/${tag}_enc.err
for t in "3 text" "3 mixed" "1 text"; do set -- $t
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_enc_$2_l$1_launches.csv \
    python tools/time_compress.py 256 65536 $1 1 $2 > gpurun_out/${tag}_enc_$2_l$1_ncu.log 2>&1
done
python - <<PY
import csv,io,collections
for t in ("text_l3","mixed_l3","text_l1"):
    lines=[l for l in open("gpurun_out/${tag}_enc_%s_launches.csv"%t) if l.startswith(chr(34))]
    agg=collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        n=r["Kernel Name"].split("(")[0]
        a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r["Metric Value"])/1e6
    print(t, {k:(v[0],round(v[1],3)) for k,v in agg.items() if v[1]>0.05})
PY
