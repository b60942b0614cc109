mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_net.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t1.log; cat gpurun_out/t1.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-breakdown > gpurun_out/b.log 2>&1
