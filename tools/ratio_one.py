import sys, json, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, refzra, zra_b200
from zra_b200 import synth
n=32<<20; fs=int(sys.argv[1]); lvl=int(sys.argv[2])
for name,data in (("text",synth.text(n,seed=3)),("mixed",synth.mixed(n,period=65536,seed=3))):
    ref=refzra.ref_compress_mt(data,lvl,fs,True).size
    z=zra_b200.CompressBuffer(data,lvl,fs,True)
    print(json.dumps({"data":name,"frame":fs,"level":lvl,"delta":round(z.size/ref-1,4),"env":{k:v for k,v in os.environ.items() if k.startswith("ZRA_B200")}}))
