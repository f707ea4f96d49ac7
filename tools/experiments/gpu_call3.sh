#!/bin/bash
# GPU call: encoder tests + timing + launch list
tag=${1:-r01f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encode.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
for k in mixed text; do for lv in 3 1; do
  timeout 300 python tools/time_compress.py 256 65536 $lv 3 $k >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err
done; done
timeout 300 python tools/time_compress.py 256 16384 3 3 text >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err
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
