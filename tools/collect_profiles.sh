#!/bin/bash
# Copies the text summaries of a round's evidence run (tools/gpu_evidence.sh) from gpurun_out/ into profiles/.
R=${1:-r02}
for f in gpurun_out/${R}_bench_n1.json gpurun_out/${R}_bench_reference.json gpurun_out/${R}_bench_vgg.json gpurun_out/${R}_bench_resnet_n1.json \
         gpurun_out/${R}_bench_n2*.json gpurun_out/${R}_bench_n4*.json gpurun_out/${R}_bench_n8*.json gpurun_out/${R}_bench_resnet_n4.json \
         gpurun_out/${R}_tc_peak.json gpurun_out/${R}_ncu_alexnet.md gpurun_out/${R}_ncu_bn.md gpurun_out/${R}_ncu_vgg.md \
         gpurun_out/${R}_ncu_*_stalls.txt gpurun_out/${R}_hot_*.txt gpurun_out/${R}_memcheck.log gpurun_out/${R}_racecheck.log \
         gpurun_out/${R}_launches.csv gpurun_out/${R}_vgg_parity.log; do
  [ -f "$f" ] && cp "$f" profiles/
done
[ -f gpurun_out/${R}_launches.csv ] && python tools/launch_list.py gpurun_out/${R}_launches.csv > profiles/${R}_launch_list.md
[ -f gpurun_out/${R}_ncu_traffic.json ] && cp gpurun_out/${R}_ncu_traffic.json profiles/ncu_traffic.json
ls profiles | grep ${R}_ | head -60
