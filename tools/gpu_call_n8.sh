#!/bin/bash
# N=8: weak-scaling line at 256/GPU, BASELINE config 4 (128/GPU, global batch 1024), and the 256/GPU line without the
# per-rank NUMA binding (e2e A/B)
R=${1:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${R}_topo_n8.txt 2>&1
for d in /sys/bus/pci/devices/*; do [ -f $d/numa_node ] && echo "$(basename $d) $(cat $d/class) $(cat $d/numa_node)"; done | grep " 0x0302" > gpurun_out/${R}_numa_n8.txt 2>&1
lscpu | grep -i "numa\|socket\|model name" >> gpurun_out/${R}_numa_n8.txt 2>&1
run() { timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 "$@"; }
run --steps 300 --warmup 5 > gpurun_out/${R}_bench_n8.json 2> gpurun_out/${R}_bench_n8.err; echo "rc=$?"
run --steps 300 --warmup 5 --batch 128 --no-breakdown > gpurun_out/${R}_bench_n8_b128.json 2> gpurun_out/${R}_bench_n8_b128.err; echo "rc=$?"
run --steps 300 --warmup 5 --no-breakdown --no-dp-check --no-numa > gpurun_out/${R}_bench_n8_nonuma.json 2> gpurun_out/${R}_bench_n8_nonuma.err; echo "rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${R}_bench_n8*.json")):
    try:
        d = json.load(open(f))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step")}, "dp", (d.get("dp_check") or {}).get("params_rel"), "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(f, "failed", e)
PY
cat gpurun_out/${R}_numa_n8.txt | head -12
