"""Data-parallel train step: the batch is sharded in contiguous slices across ranks,
every rank computes dW_local = (1/B_global) * sum over its slice straight into one flat fp32
gradient slab (checkpoint order + the summed log-likelihood in the tail slot), ONE all-reduce
(NCCL over NVLink on GPUs, gloo in the CPU tests) sums the slabs, and every rank applies the
identical SGD step, so replicas stay in lock-step without ever exchanging parameters.

torch.distributed is plumbing only; the compute engine is anything with
    fwd_bwd(x, labels, grad_scale) / grad_slab() -> 1-D torch tensor [P+1] / update(lr)
-- cnn_b200.api.Net on a GPU (NetEngine below), an oracle-backed stand-in in tests/.
BatchNorm statistics are per rank (the reference at B_local); SyncBN is future work.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(global_batch, world, rank):
    """Contiguous slice [first, first+count) of the global batch owned by `rank`."""
    assert global_batch % world == 0, "global batch must divide evenly (the reference has no ragged batches)"
    count = global_batch // world
    return rank * count, count


class NetEngine:
    """Adapter: cnn_b200.api.Net -> the engine protocol."""

    def __init__(self, net):
        self.net = net
        self._slab = net.grad_slab()

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
