mkdir -p gpurun_out
(timeout 500 python -m pytest tests/test_gpu_dist.py -m gpu -x -q 2>&1 | tail -12) > gpurun_out/t_dist.log; tail -3 gpurun_out/t_dist.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$? lines=$(wc -l < gpurun_out/bench_n2.json)"; cut -c1-120 gpurun_out/bench_n2.json
CNN_DBG_NOAROVERLAP=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown 2>/dev/null | cut -c1-120
