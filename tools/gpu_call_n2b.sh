#!/bin/bash
for flag in "" "--no-peer"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 300 --warmup 5 --batch 128 $flag --no-breakdown 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('flag=[$flag] ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['dp_check'] and (d['dp_check']['params_rel'], d['dp_check']['replicas_bit_identical']))"
done
python bench.py --steps 300 --warmup 5 --batch 128 --no-breakdown --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 B=128 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
