mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -8) > gpurun_out/t_dist.log; cat gpurun_out/t_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | cut -c1-900
(time timeout 300 python bench.py --impl reference --steps 4 --warmup 1) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
