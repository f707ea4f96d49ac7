#!/bin/bash
# GPU call: encoder kernel breakdown (ncu launch list) for L3 / L1 on mixed and text data.
tag=${1:-r01e}
mkdir -p gpurun_out
for lv in 3 1; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_enc_l${lv}_launches.csv \
    python tools/time_compress.py 256 65536 $lv 1 mixed > gpurun_out/${tag}_enc_l${lv}_ncu.log 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_enc_text_l3_launches.csv \
    python tools/time_compress.py 256 65536 3 1 text > gpurun_out/${tag}_enc_text_l3_ncu.log 2>&1
for k in mixed text; do for lv in 3 1; do
  timeout 300 python tools/time_compress.py 256 65536 $lv 3 $k >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err
done; done
timeout 300 python tools/time_compress.py 256 16384 3 3 text >> gpurun_out/${tag}_enc.jsonl 2>> gpurun_out/${tag}_enc.err
cat gpurun_out/${tag}_enc.jsonl
python - <<PY
import csv,io,collections
for t in ("l3","l1","text_l3"):
    lines=[l for l in open("gpurun_out/${tag}_enc_%s_launches.csv"%t) if l.startswith(chr(34))]
    agg=collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        n=r["Kernel Name"].split("(")[0]
        a=agg.setdefault(n,[0,0.0]); a[0]+=1; a[1]+=float(r["Metric Value"])/1e6
    print(t, {k:(v[0],round(v[1],3)) for k,v in agg.items()})
PY
