mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'thin_fwd' --launch-skip 4 -c 1 -o gpurun_out/r01_thinfused python bench.py --steps 1 --warmup 3 --no-cpu --no-breakdown > gpurun_out/ncu_tf.log 2>&1
tail -2 gpurun_out/ncu_tf.log
