#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/r02_bench6.json 2> gpurun_out/r02_bench6.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench6.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')})
bd={k:v for k,v in d['breakdown'].items() if isinstance(v,dict)}
for k,v in sorted(bd.items(), key=lambda kv:-kv[1]['us']): print(k, v['us'])
print(d['breakdown']['_total_us_eager_with_event_gaps'])
PY
ncu --set full --clock-control none --import-source on -k regex:'s2_' -s 20 -c 15 -o gpurun_out/r02_s2 python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/r02_ncu_s2.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_s2.ncu-rep > gpurun_out/r02_ncu_s2.md
python tools/ncu_stalls.py gpurun_out/r02_s2.ncu-rep > gpurun_out/r02_ncu_s2_stalls.txt
for k in "s2_gemm_kernel<0>" "s2_gemm_kernel<1>" s2_wgrad_kernel s2_pack_d; do echo "== $k"; python tools/ncu_hot.py gpurun_out/r02_s2.ncu-rep "$k" 12; done > gpurun_out/r02_s2_hot.txt 2>&1
rm -f gpurun_out/r02_s2.ncu-rep
cat gpurun_out/r02_ncu_s2.md | cut -c1-260
cat gpurun_out/r02_ncu_s2_stalls.txt
