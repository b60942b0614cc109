#!/bin/bash
# Evidence set of a round (one B200): tests, micro-benchmark, bench lines, ncu launch list + full captures.
# Outputs land in gpurun_out/; tools/collect_profiles.sh copies the summaries into profiles/.
R=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${R}_pytest_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${R}_pytest_final.log; tail -4 gpurun_out/${R}_pytest_final.log
tools/_build/tc_peak > gpurun_out/${R}_tc_peak.json 2>&1; cat gpurun_out/${R}_tc_peak.json
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_step.py > gpurun_out/${R}_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/${R}_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/memcheck_step.py > gpurun_out/${R}_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/${R}_racecheck.log
# ncu first (its traffic file feeds the bench line of the same binary)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${R}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -s 60 -c 40 -o gpurun_out/${R}_alexnet python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/${R}_ncu_alexnet.log 2>&1
ncu --set full --clock-control none -k regex:'bn_|relu_|sgd_|softmax|linear' -c 24 -o gpurun_out/${R}_bn python bench.py --bn --batch 64 --steps 1 --warmup 3 --no-cpu --no-breakdown > gpurun_out/${R}_ncu_bn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'s1_' -s 30 -c 14 -o gpurun_out/${R}_vgg python bench.py --net vgg_style --batch 32 --steps 1 --warmup 3 --no-cpu --no-breakdown > gpurun_out/${R}_ncu_vgg.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/ncu_traffic.py gpurun_out/${R}_alexnet.ncu-rep > profiles/ncu_traffic.json 2>/dev/null; head -c 600 profiles/ncu_traffic.json
cp profiles/ncu_traffic.json gpurun_out/${R}_ncu_traffic.json
# the reports themselves are too big to travel back (64 MiB limit): condense them here, keep the text
for n in alexnet bn vgg; do
  python tools/ncu_summary.py gpurun_out/${R}_$n.ncu-rep > gpurun_out/${R}_ncu_$n.md 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/${R}_$n.ncu-rep > gpurun_out/${R}_ncu_${n}_stalls.txt 2>/dev/null
done
for k in head_fwd head_wgrad s2_gemm s2_wgrad; do python tools/ncu_hot.py gpurun_out/${R}_alexnet.ncu-rep $k 25 > gpurun_out/${R}_hot_$k.txt 2>/dev/null; done
for k in s1_gemm s1_wgrad; do python tools/ncu_hot.py gpurun_out/${R}_vgg.ncu-rep $k 25 > gpurun_out/${R}_hot_$k.txt 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 200 --warmup 5 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 2 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/${R}_bench_reference.err; echo "ref rc=$?"
python bench.py --net vgg_style --batch 128 --steps 5 --warmup 3 --no-cpu > gpurun_out/${R}_bench_vgg.json 2> gpurun_out/${R}_bench_vgg.err; echo "vgg rc=$?"
python bench.py --net resnet18_shaped --batch 128 --precision bf16 --steps 5 --warmup 3 --no-cpu > gpurun_out/${R}_bench_resnet_n1.json 2> gpurun_out/${R}_bench_resnet_n1.err; echo "resnet rc=$?"
python - <<PY
import json
for f in ("bench_n1", "bench_vgg", "bench_resnet_n1", "bench_reference"):
    try:
        d = json.load(open("gpurun_out/${R}_%s.json" % f))
        print(f, {k: d.get(k) for k in ("value", "ms_per_step", "roofline", "north_star_pair")}, (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
