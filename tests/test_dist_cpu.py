"""world_size-2 gloo test of the data-parallel step (cnn_b200/dist.py) on CPU.  The compute engine
is an oracle-backed stand-in, so this checks the host logic: sharding, 1/B_global scaling, the
single all-reduce of the gradient slab (loss in the tail slot), and replicated SGD -- against the
single-process oracle at the global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN

B_GLOBAL, STEPS, LR = 4, 2, 1e-3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleEngine:
    """Engine protocol on top of oracle.port.Net (which folds SGD into train_step): run with
    lr = 0 to obtain gradients (scaled 1/B_local by the reference), rescale to 1/B_global."""

    def __init__(self, spec, b_local, params):
        from oracle import port
        self.net = port.Net(spec, b_local, 3, 224, 224)
        self.net.set_params(params)
        self.b_local = b_local
        self._slab = torch.zeros(self.net.n_params + 1, dtype=torch.float32)

    def fwd_bwd(self, x, labels, grad_scale):
        loss, _, _ = self.net.train_step(x, labels, 0.0)
        g = self.net.get_grads() * np.float32(self.b_local * grad_scale)
        self._slab[:-1] = torch.from_numpy(g)
        self._slab[-1] = float(-loss * self.b_local)  # sum_b log p[label]

    def grad_slab(self):
        return self._slab

    def update(self, lr):
        p = self.net.get_params() - np.float32(lr) * self._slab[:-1].numpy()
        self.net.set_params(p)


def _worker(rank, world, port_no, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cnn_b200.dist import dp_train_step, shard_range, broadcast_params
    from cnn_b200.nets import alexnet_lite
    from cnn_b200.synth import synth_images, synth_labels
    first, count = shard_range(B_GLOBAL, world, rank)
    init = np.fromfile(os.path.join(GOLDEN, "alexnet_init.model"), np.float32)
    params = broadcast_params(init if rank == 0 else np.zeros_like(init))
    assert np.array_equal(params, init)
    eng = OracleEngine(alexnet_lite(3), count, params)
    x = synth_images(count, seed=1234, first_image=first)
    lab = synth_labels(count, 3, first_image=first)
    losses = []
    for _ in range(STEPS):
        losses.append(float(dp_train_step(eng, x, lab, LR, B_GLOBAL)))
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), params=eng.net.get_params(), losses=np.array(losses))
    dist.destroy_process_group()


def test_shard_range():
    from cnn_b200.dist import shard_range
    assert [shard_range(1024, 8, r) for r in (0, 3, 7)] == [(0, 128), (384, 128), (896, 128)]
    with pytest.raises(AssertionError):
        shard_range(10, 4, 0)


def test_two_rank_dp_matches_single_process_reference(tmp_path, train_golden):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(tmp_path / f"rank{r}.npz") for r in range(world))
    assert np.array_equal(r0["params"], r1["params"])            # replicas stay bit-identical
    assert np.array_equal(r0["losses"], r1["losses"])
    # the reference's own trajectory at B=4 (tests/golden/train_golden.npz, produced by the reference)
    g = train_golden
    for s in range(STEPS):
        assert abs(r0["losses"][s] - g[f"loss{s}"]) <= 1e-5 * max(1.0, abs(g[f"loss{s}"]))
    from oracle import port
    from cnn_b200.nets import alexnet_lite
    from cnn_b200.synth import synth_images, synth_labels
    single = port.Net(alexnet_lite(3), B_GLOBAL, 3, 224, 224)
    single.set_params(np.fromfile(os.path.join(GOLDEN, "alexnet_init.model"), np.float32))
    x, lab = synth_images(B_GLOBAL), synth_labels(B_GLOBAL)
    for _ in range(STEPS):
        single.train_step(x, lab, LR)
    ref = single.get_params()
    err = np.abs(r0["params"] - ref).max() / np.abs(ref).max()
    assert err <= 1e-6, err
