#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_net.py -m gpu -x -q > gpurun_out/r02_pytest3.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest3.log
tail -8 gpurun_out/r02_pytest3.log
python bench.py --steps 50 --warmup 5 --no-cpu --no-breakdown > gpurun_out/r02_bench_lazy2.json 2> gpurun_out/r02_bench_lazy2.err
python -c "import json; d=json.load(open('gpurun_out/r02_bench_lazy2.json')); print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e'].get('u8_images'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches2.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > /dev/null 2>&1
python tools/launch_list.py gpurun_out/r02_launches2.csv | head -12
ncu --set full --clock-control none --import-source on -k regex:head_ -s 4 -c 2 -o gpurun_out/r02_head2 python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/r02_ncu_head2.log 2>&1
ls -la gpurun_out/r02_head2.ncu-rep
