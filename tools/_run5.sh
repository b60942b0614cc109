mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fused" 2>&1 | tail -3) > gpurun_out/t1.log; cat gpurun_out/t1.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu --no-breakdown | cut -c1-220
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/launches_s.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/b.log 2>&1
