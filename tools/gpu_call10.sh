#!/bin/bash
python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -x -q -k "conv_vs_oracle or conv_golden or trajectory or full_batch or lazy or lr_schedule" 2>&1 | tail -3
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02_bench10.json 2> gpurun_out/r02_bench10.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench10.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
bd={k:v for k,v in d['breakdown'].items() if isinstance(v,dict)}
for k,v in sorted(bd.items(), key=lambda kv:-kv[1]['us'])[:24]: print(k, v['us'])
PY
