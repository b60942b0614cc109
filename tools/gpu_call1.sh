#!/bin/bash
# first GPU call of round 2: tests, lazy vs materialising step time, ncu of the new head kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest1.log
tail -15 gpurun_out/r02_pytest1.log
python bench.py --steps 50 --warmup 5 --no-cpu --no-breakdown > gpurun_out/r02_bench_lazy.json 2> gpurun_out/r02_bench_lazy.err
CNN_LAZY_HEAD=0 python bench.py --steps 50 --warmup 5 --no-cpu --no-breakdown > gpurun_out/r02_bench_full.json 2> gpurun_out/r02_bench_full.err
for t in 3 5; do CNN_HEAD_TRP=$t python bench.py --steps 50 --warmup 5 --no-cpu --no-breakdown 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('TRP=$t', d['ms_per_step'])"; done
for t in 1 3 4; do CNN_HEADWG_TRP=$t python bench.py --steps 50 --warmup 5 --no-cpu --no-breakdown 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('WGTRP=$t', d['ms_per_step'])"; done
CNN_HEAD_NBUF=3 CNN_HEAD_TRP=2 python bench.py --steps 50 --warmup 5 --no-cpu --no-breakdown 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NBUF3 TRP2', d['ms_per_step'])"
python -c "import json; [print(f, json.load(open('gpurun_out/'+f))['ms_per_step']) for f in ('r02_bench_lazy.json','r02_bench_full.json')]"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches1.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > /dev/null 2>&1
python tools/launch_list.py gpurun_out/r02_launches1.csv | head -30
ncu --set full --clock-control none --import-source on -k regex:head_ -s 4 -c 4 -o gpurun_out/r02_head python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/r02_ncu_head.log 2>&1
ls -la gpurun_out/r02_head.ncu-rep
