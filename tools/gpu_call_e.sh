#!/bin/bash
# full GPU suite + the bench line (both arms). usage: tools/gpu_call_e.sh <tag>
tag=$1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -4 gpurun_out/${tag}_pytest.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
cat gpurun_out/${tag}_bench_ref.json | cut -c1-400
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${tag}_bench.err
python - <<'PY'
import json,sys
try:
    l=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json').read().strip().splitlines()[-1]) if False else None
except Exception as e: print(e)
PY
python -c "
import json
l=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['whole_step_frac'], l['roofline']['kernels_ms_per_step'])
print('ra', l['random_access']['value'], l['random_access']['e2e']['value'])
c=l['compress']; print('compress', {k:(v['value'], v.get('e2e',{}).get('value'), v.get('ratio_vs_reference',{}).get('size_delta')) for k,v in c['levels'].items()}, c.get('frames_256KiB_L3'))
s=l['streaming']; print('streaming', s['value'], s.get('device_resident_decode_256KiB_frames'))
print('cpu', l.get('cpu_baseline',{}).get('value'))
"
