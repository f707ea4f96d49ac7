#!/bin/bash
# --gpus 8 call: the box's aggregate PCIe ceiling with 8 ranks copying at once, then the N=8 bench (both arms).
tag=${1:-r03n8}
n=${2:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
free -g | head -2 > gpurun_out/${tag}_mem.txt; nproc >> gpurun_out/${tag}_mem.txt; lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/${tag}_mem.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 tools/pcie_probe_ranks.py 1024 > gpurun_out/${tag}_pcie.json 2> gpurun_out/${tag}_pcie.err; echo "pcie rc=$?"
cat gpurun_out/${tag}_pcie.json
timeout 300 python tools/pcie_probe_ranks.py 1024 > gpurun_out/${tag}_pcie1.json 2>> gpurun_out/${tag}_pcie.err; cat gpurun_out/${tag}_pcie1.json
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/${tag}_bench.err | cut -c1-300
python -c "
import json
l=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', l['e2e']['value'], l.get('gather'))
print('ra', l['random_access']['value'], l['random_access']['e2e']['value'], l['random_access'].get('sharded',{}).get('value'), l['random_access'].get('sharded',{}).get('reads_per_step'))
c=l['compress']; print('compress', c.get('error'), {k:v['value'] for k,v in c.get('levels',{}).items()}, c.get('sharded_archive_verified'))
print('streaming', l['streaming'].get('value'))
"
cat gpurun_out/${tag}_mem.txt
