#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_pytest_n2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_n2.log
tail -15 gpurun_out/r02_pytest_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 200 --warmup 5 --batch 128 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; echo "bench n2 rc=$?"; tail -5 gpurun_out/r02_bench_n2b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n2b.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','dp_check','roofline')})
print(d['e2e'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 200 --warmup 5 --batch 128 --no-peer --no-dp-check --no-breakdown 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('NCCL path: ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
