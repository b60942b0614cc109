#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/fullstep_parity_vgg.py --batch 16 > gpurun_out/r02_vgg_parity4.log 2>&1; echo "parity rc=$?"; tail -26 gpurun_out/r02_vgg_parity4.log
timeout 600 python bench.py --net vgg_style --batch 128 --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_vgg3.json 2> gpurun_out/r02_bench_vgg3.err; echo "vgg rc=$?"; tail -3 gpurun_out/r02_bench_vgg3.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/r02_bench_vgg3.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','roofline')})
    for k,v in sorted(d['breakdown'].items(), key=lambda kv:-(kv[1]['us'] if isinstance(kv[1],dict) else 0))[:40]: print(k,v)
except Exception as e: print('vgg parse failed', e)
PY
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest6.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest6.log
tail -6 gpurun_out/r02_pytest6.log
