#!/usr/bin/env python
"""Run under torchrun with N ranks: N-GPU data-parallel trajectory == 1-GPU trajectory at the same
global batch (SURVEY §4 test plan iv).  Rank 0 prints DP_CHECK OK / FAILED."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from cnn_b200.api import Context, Net
from cnn_b200.dist import NetEngine, dp_train_step, init_native_dist, shard_range
from cnn_b200.nets import alexnet_lite, insert_bn_params
from cnn_b200.synth import synth_images, synth_labels


def main():
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    Bg, steps, lr = 32, 3, 1e-3
    first, count = shard_range(Bg, world, rank)
    init = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "alexnet_init.model"), np.float32)
    ctx = Context(local)
    bn = "--bn" in sys.argv              # BatchNorm net with SyncBN: N ranks at B/N == one GPU at B
    spec = alexnet_lite(3, batch_norm=bn)
    if bn:
        init = insert_bn_params(spec, init)
    net = Net(ctx, spec, count)
    net.set_params(init)
    native = "--native" in sys.argv or bn      # all-reduce issued by the library inside the step graph (dist.cu)
    if native:
        init_native_dist(ctx)
    peer = "--peer" in sys.argv
    if peer:
        assert native, "--peer needs --native"
        ok_peer = net.enable_peer_exchange()
        if rank == 0:
            print("peer exchange enabled:", ok_peer)
    if bn:
        from cnn_b200._lib import check
        check(ctx.L.cnn_dist_set_sync_bn(ctx._h, 1), "cnn_dist_set_sync_bn")
    eng = NetEngine(net, native_dist=native)
    seed = int(sys.argv[sys.argv.index("--seed") + 1]) if "--seed" in sys.argv else 1234
    x = ctx.to_device(synth_images(count, seed=seed, first_image=first))
    lab = ctx.to_device(synth_labels(count, 3, first_image=first), torch.int32)
    losses = []
    for _ in range(steps):
        losses.append(float(dp_train_step(eng, x, lab, lr, Bg).item()))
    params = net.get_params()
    ok = True
    if rank == 0:
        if bn:  # the single-GPU reference run must not all-reduce its statistics
            from cnn_b200._lib import check
            check(ctx.L.cnn_dist_set_sync_bn(ctx._h, 0), "cnn_dist_set_sync_bn")
        ref = Net(ctx, spec, Bg)
        ref.set_params(init)
        xr = ctx.to_device(synth_images(Bg, seed=seed))
        lr_ = ctx.to_device(synth_labels(Bg, 3), torch.int32)
        rl = []
        for _ in range(steps):
            ref.train_step(xr, lr_, lr)
            ctx.sync()
            rl.append(float(ref.loss_from_slab()))
        rp = ref.get_params()
        e_p = float(np.abs(params - rp).max() / np.abs(rp).max())
        from cnn_b200.nets import param_layout
        for li, kind, off, n in param_layout(spec)[0]:
            d = float(np.abs(params[off:off + n] - rp[off:off + n]).max())
            print(f"   layer {li} {kind}: max abs diff {d:.3e} (max abs {float(np.abs(rp[off:off + n]).max()):.3e})")
        e_l = max(abs(a - b) / max(1.0, abs(b)) for a, b in zip(losses, rl))
        ok = e_p <= 1e-4 and e_l <= 1e-4
        print(f"world {world} ({'library NCCL, in-graph' if native else 'torch.distributed'}{', SyncBN' if bn else ''}): losses {losses} vs 1-GPU {rl}; rel.err params {e_p:.2e} loss {e_l:.2e}")
    # replicas identical?
    t = torch.from_numpy(params).cuda()
    mx, mn = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN)
    same = bool(torch.equal(mx, mn))
    if rank == 0:
        print("replicas bit-identical:", same)
        print("DP_CHECK OK" if ok and same else "DP_CHECK FAILED")
    net.close()
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
