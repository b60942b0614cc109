#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'s1_gemm|s1_wgrad_kernel' -s 6 -c 6 -o gpurun_out/r02_s1 python bench.py --net vgg_style --batch 32 --steps 1 --warmup 3 --no-cpu --no-breakdown > gpurun_out/r02_ncu_s1.log 2>&1
python tools/ncu_summary.py gpurun_out/r02_s1.ncu-rep > gpurun_out/r02_ncu_s1.md
python tools/ncu_stalls.py gpurun_out/r02_s1.ncu-rep > gpurun_out/r02_ncu_s1_stalls.txt
for k in "s1_gemm_kernel<0>" "s1_gemm_kernel<1>" s1_wgrad_kernel; do echo "== $k"; python tools/ncu_hot.py gpurun_out/r02_s1.ncu-rep "$k" 14; done > gpurun_out/r02_s1_hot.txt 2>&1
rm -f gpurun_out/r02_s1.ncu-rep
cat gpurun_out/r02_ncu_s1.md | cut -c1-250
cat gpurun_out/r02_ncu_s1_stalls.txt
cat gpurun_out/r02_s1_hot.txt
