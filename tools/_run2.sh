mkdir -p gpurun_out
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/b.log 2>&1
timeout 500 ncu --set full --import-source on --clock-control none -k regex:'thin_fwd|thin_dgrad|thin_wgrad_kernel|band|s2_gemm|s2_wgrad_kernel|s2_pack_x|s2_pack_d' --launch-skip 60 -c 20 -o gpurun_out/r01_full python bench.py --steps 1 --warmup 3 --no-cpu --no-breakdown > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
