"""Multi-GPU data parallelism on real devices (skipped on a 1-GPU box): N-rank NCCL trajectory equals
the single-GPU trajectory at the same global batch, replicas stay bit-identical, and bench.py's
torchrun contract prints its JSON line."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(n, script_args, port):
    return subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                           "--master-addr", "127.0.0.1", "--master-port", str(port)] + script_args,
                          capture_output=True, text=True, timeout=600, cwd=ROOT)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_trajectory_matches_single_gpu():
    n = min(torch.cuda.device_count(), 4)
    r = _torchrun(n, ["tools/dp_check.py"], 29511)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_trajectory_with_library_allreduce_in_graph():
    """cnn_dist_init + cnn_net_train_step(do_update=3): the slab all-reduce is issued by the library on
    its own stream inside the step's CUDA graph; same trajectory, replicas bit-identical."""
    n = min(torch.cuda.device_count(), 4)
    r = _torchrun(n, ["tools/dp_check.py", "--native"], 29513)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout and "library NCCL" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_dp_trajectory_with_peer_memory_exchange():
    """cnn_net_enable_peer_exchange: the gradient sum over ranks and the SGD step as ONE kernel over NVLink peer
    memory (every rank reads all slabs, rank-order sum); same trajectory as one GPU, replicas bit-identical."""
    n = min(torch.cuda.device_count(), 4)
    r = _torchrun(n, ["tools/dp_check.py", "--native", "--peer"], 29515)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout and "peer exchange enabled: True" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_sync_batchnorm_matches_single_gpu_global_batch():
    """BatchNorm net under data parallelism with cnn_dist_set_sync_bn: the batch statistics and backward
    sums are all-reduced, so N ranks at B/N walk the trajectory of one GPU at B -- which is the
    single-process reference at B_global (batchnorm2d.cpp:46-61, :118-147)."""
    n = min(torch.cuda.device_count(), 4)
    r = _torchrun(n, ["tools/dp_check.py", "--bn"], 29514)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0 and "DP_CHECK OK" in r.stdout and "SyncBN" in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_bench_contract_multi_gpu():
    r = _torchrun(2, ["bench.py", "--gpus", "2", "--steps", "3", "--warmup", "3", "--batch", "64", "--no-breakdown"], 29512)
    print(r.stdout[-2000:], r.stderr[-3000:])
    assert r.returncode == 0
    line = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["n_gpus"] == 2 and line["scaling"] == "weak" and line["value"] > 0 and line["gpu_launches"] > 0
    dc = line["dp_check"]
    assert dc["loss_rel"] <= 1e-4 and dc["params_rel"] <= 1e-4 and dc["replicas_bit_identical"]
    assert dc["sync_bn"]["loss_rel"] <= 1e-4 and dc["sync_bn"]["params_rel"] <= 1e-4
