#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest4.log
tail -6 gpurun_out/r02_pytest4.log
timeout 300 python bench.py --steps 50 --warmup 5 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench3.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench3.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','materialized','roofline','north_star_pair')})
print(d['e2e'])
for k,v in sorted(d['breakdown'].items(), key=lambda kv:-(kv[1]['us'] if isinstance(kv[1],dict) else 0))[:14]: print(k,v)
PY
timeout 600 python bench.py --net vgg_style --batch 128 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_vgg1.json 2> gpurun_out/r02_bench_vgg1.err; echo "vgg rc=$?"; tail -3 gpurun_out/r02_bench_vgg1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_vgg1.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','roofline')})
    for k,v in sorted(d['breakdown'].items(), key=lambda kv:-(kv[1]['us'] if isinstance(kv[1],dict) else 0))[:30]: print(k,v)
except Exception as e: print('vgg parse failed', e)
PY
timeout 600 python bench.py --net resnet18_shaped --batch 128 --precision bf16 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_rn1.json 2> gpurun_out/r02_bench_rn1.err; echo "resnet rc=$?"; tail -3 gpurun_out/r02_bench_rn1.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_rn1.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','roofline')})
    for k,v in sorted(d['breakdown'].items(), key=lambda kv:-(kv[1]['us'] if isinstance(kv[1],dict) else 0))[:12]: print(k,v)
except Exception as e: print('resnet parse failed', e)
PY
