#!/bin/bash
# Final check of a round with little GPU time left: the whole GPU test suite, then the three single-GPU bench lines.
R=${1:-r02}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_final.log; tail -3 gpurun_out/${R}_pytest_final.log
timeout 150 python bench.py --steps 300 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
timeout 100 python bench.py --net vgg_style --batch 128 --steps 5 --warmup 3 --no-cpu > gpurun_out/${R}_bench_vgg.json 2> gpurun_out/${R}_bench_vgg.err; echo "vgg rc=$?"
timeout 100 python bench.py --net resnet18_shaped --batch 128 --precision bf16 --bn --steps 10 --warmup 3 --no-cpu > gpurun_out/${R}_bench_resnet_n1.json 2> gpurun_out/${R}_bench_resnet_n1.err; echo "resnet rc=$?"
python - <<PY
import json
for n in ("bench_n1", "bench_vgg", "bench_resnet_n1"):
    try:
        d = json.load(open("gpurun_out/${R}_%s.json" % n))
        print(n, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["kernel"], d["roofline"]["frac"], d["clocks"]["reasons"])
    except Exception as e:
        print(n, "failed", e)
PY
