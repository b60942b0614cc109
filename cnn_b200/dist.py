"""Data-parallel train step: the batch is sharded in contiguous slices across ranks,
every rank computes dW_local = (1/B_global) * sum over its slice straight into one flat fp32
gradient slab (checkpoint order + the summed log-likelihood in the tail slot), ONE all-reduce
(NCCL over NVLink on GPUs, gloo in the CPU tests) sums the slabs, and every rank applies the
identical SGD step, so replicas stay in lock-step without ever exchanging parameters.

torch.distributed is plumbing only; the compute engine is anything with
    fwd_bwd(x, labels, grad_scale) / grad_slab() -> 1-D torch tensor [P+1] / update(lr)
-- cnn_b200.api.Net on a GPU (NetEngine below), an oracle-backed stand-in in tests/.
BatchNorm statistics are per rank by default (the reference at B_local); cnn_dist_set_sync_bn makes the library
all-reduce them (bn.cu), so N ranks at B/N reproduce the single-process reference at B.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(global_batch, world, rank):
    """Contiguous slice [first, first+count) of the global batch owned by `rank`."""
    assert global_batch % world == 0, "global batch must divide evenly (the reference has no ragged batches)"
    count = global_batch // world
    return rank * count, count


def init_native_dist(ctx, group=None):
    """Give this rank's library context its own NCCL communicator (cnn_dist_init, dist.cu): rank 0's
    ncclUniqueId travels through torch.distributed, which stays pure plumbing.  Afterwards the
    all-reduce of the gradient slab is issued by the library on its own stream, inside the step's CUDA
    graph (cnn_net_train_step with do_update = 3)."""
    import ctypes as C
    from ._lib import check
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    raw = (C.c_ubyte * 128)()
    if rank == 0:
        check(ctx.L.cnn_dist_unique_id(raw), "cnn_dist_unique_id")
    t = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.to(ctx.device)
    dist.broadcast(t, src=0, group=group)
    raw = (C.c_ubyte * 128)(*t.cpu().tolist())
    check(ctx.L.cnn_dist_init(ctx._h, rank, world, raw), "cnn_dist_init")
    return world


class NetEngine:
    """Adapter: cnn_b200.api.Net -> the engine protocol.  native_dist: the library owns an NCCL
    communicator (init_native_dist) and runs forward + backward + all-reduce + SGD as one graph."""

    def __init__(self, net, native_dist=False):
        self.net = net
        self._slab = net.grad_slab()
        self.native_dist = native_dist

    def step_fused(self, x, labels, lr, global_batch):
        self.net.train_step(x, labels, lr, grad_scale=1.0 / global_batch, do_update=3)

    def fwd_bwd(self, x, labels, grad_scale):
        self.net.train_step(x, labels, 0.0, grad_scale=grad_scale, do_update=False)

    def grad_slab(self):
        return self._slab

    def update(self, lr):
        self.net.update(lr)

    @property
    def stream(self):
        return self.net.ctx.stream


def dp_train_step(engine, x_shard, labels_shard, lr, global_batch, group=None):
    """One data-parallel step on this rank's shard.  Returns the loss tensor (-sum log p / B_global),
    left on the slab's device so the caller decides when to synchronise."""
    if getattr(engine, "native_dist", False):
        # one CUDA-graph launch per rank: fwd + bwd + ncclAllReduce(slab) + SGD, all on the engine's stream
        engine.step_fused(x_shard, labels_shard, lr, global_batch)
        slab, stream = engine.grad_slab(), engine.stream
        with torch.cuda.stream(stream):
            loss = slab[-1] * (-1.0 / global_batch)
        loss.record_stream(stream)
        torch.cuda.current_stream().wait_stream(stream)
        return loss
    engine.fwd_bwd(x_shard, labels_shard, 1.0 / global_batch)
    slab = engine.grad_slab()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world > 1:
        stream = getattr(engine, "stream", None)
        if stream is not None:
            with torch.cuda.stream(stream):
                dist.all_reduce(slab, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.all_reduce(slab, op=dist.ReduceOp.SUM, group=group)
    engine.update(lr)
    stream = getattr(engine, "stream", None)
    if stream is not None:  # read the tail slot in stream order (after the all-reduce, before the next step)
        with torch.cuda.stream(stream):
            loss = slab[-1] * (-1.0 / global_batch)
        loss.record_stream(stream)
        torch.cuda.current_stream().wait_stream(stream)
        return loss
    return slab[-1] * (-1.0 / global_batch)


def broadcast_params(flat_params, src=0, group=None):
    """Rank `src`'s parameter vector (numpy) on every rank: replicas must start identical."""
    t = torch.from_numpy(np.ascontiguousarray(flat_params, np.float32))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if dist.get_backend(group) == "nccl":
            t = t.cuda()
        dist.broadcast(t, src=src, group=group)
    return t.cpu().numpy()
