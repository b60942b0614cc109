#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest8.log 2>&1; tail -4 gpurun_out/r02_pytest8.log
timeout 600 python bench.py --net resnet18_shaped --batch 128 --precision bf16 --steps 5 --warmup 3 --no-cpu > gpurun_out/r02_bench_rn3.json 2> gpurun_out/r02_bench_rn3.err; echo "resnet rc=$?"; tail -3 gpurun_out/r02_bench_rn3.err
python - <<'PY'
import json, collections
d=json.load(open('gpurun_out/r02_bench_rn3.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
bd={k:v for k,v in d['breakdown'].items() if isinstance(v,dict)}
tot=sum(v['us'] for v in bd.values())
agg=collections.Counter()
for k,v in bd.items():
    kern=k.split(':',1)[1] if ':' in k else k
    agg[kern.split('<')[0]]+=v['us']
for k,v in agg.most_common(16): print(f"{k:40s} {v:9.1f} us {100*v/tot:5.1f}%")
for k,v in sorted(bd.items(), key=lambda kv:-kv[1]['us'])[:14]: print(k, v)
PY
python bench.py --bn --steps 50 --warmup 5 --no-cpu --no-breakdown 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('alexnet+BN B=256 ms/step', d['ms_per_step'], d['value'])"
