#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --net vgg_style --batch 128 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_vgg4.json 2> gpurun_out/r02_bench_vgg4.err; echo "vgg rc=$?"
python - <<'PY'
import json, collections
d=json.load(open('gpurun_out/r02_bench_vgg4.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','roofline')})
bd={k:v for k,v in d['breakdown'].items() if isinstance(v,dict)}
tot=sum(v['us'] for v in bd.values())
agg=collections.Counter()
for k,v in bd.items():
    kern=k.split(':',1)[1] if ':' in k else k
    agg[kern.split('<')[0]+('<'+kern.split('<')[1] if 's1_gemm' in kern else '')]+=v['us']
for k,v in agg.most_common(20): print(f"{k:40s} {v:9.1f} us {100*v/tot:5.1f}%")
for k,v in sorted(bd.items(), key=lambda kv:-kv[1]['us'])[:30]: print(k, v['us'])
PY
