#!/bin/bash
# usage: gpu_call_nN.sh N tag  -- AlexNet-lite weak scaling line at 256/GPU and 128/GPU (config 4 at N=8), resnet18-shaped at N=4
N=$1; R=${2:-r02}
mkdir -p gpurun_out
run() { timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus $N "$@"; }
run --steps 300 --warmup 5 > gpurun_out/${R}_bench_n${N}.json 2> gpurun_out/${R}_bench_n${N}.err; echo "rc=$?"
run --steps 300 --warmup 5 --batch 128 --no-breakdown > gpurun_out/${R}_bench_n${N}_b128.json 2> gpurun_out/${R}_bench_n${N}_b128.err; echo "rc=$?"
if [ "$N" = "4" ]; then
  run --net resnet18_shaped --batch 128 --precision bf16 --bn --steps 10 --warmup 3 --no-dp-check > gpurun_out/${R}_bench_resnet_n4.json 2> gpurun_out/${R}_bench_resnet_n4.err; echo "resnet rc=$?"; tail -3 gpurun_out/${R}_bench_resnet_n4.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${R}_bench*n${N}*.json")):
    try:
        d = json.load(open(f))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "dp_check")}, "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["allreduce"][:40])
    except Exception as e:
        print(f, "failed", e)
PY
