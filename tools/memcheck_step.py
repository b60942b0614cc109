#!/usr/bin/env python
"""Small workload for `compute-sanitizer --tool memcheck|racecheck`: the packed stride-2 and stride-1 kernels at odd
shapes, the band pooling kernels, BatchNorm, engine steps (eager and graph) of AlexNet-lite at B=2 -- lazy head
(with on-demand materialisation) and materialising path -- and of a small VGG-style / ResNet-shaped stack.
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_step.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cnn_b200.api import Context, Net
from cnn_b200.nets import alexnet_lite
from cnn_b200.synth import synth_images, synth_labels


def main():
    ctx = Context(0)
    rng = np.random.default_rng(0)
    for (B, Cin, H, W, Cout) in [(2, 16, 20, 18, 32), (1, 32, 9, 37, 16), (2, 48, 11, 12, 48), (1, 128, 10, 10, 128), (3, 64, 13, 13, 128)]:
        x = ctx.to_device(rng.random((B, Cin, H, W), dtype=np.float32))
        w = ctx.to_device((rng.standard_normal((Cout, Cin, 3, 3)) / 10).astype(np.float32))
        b = ctx.to_device(np.zeros(Cout, np.float32))
        y = ctx.conv2d_forward(x, w, b, 2)
        ctx.conv2d_backward(x, w, torch.randn_like(y), 2)
    for (B, Cin, H, W, Cout, k, st) in [(2, 32, 20, 18, 32, 3, 1), (1, 64, 9, 37, 96, 3, 1), (3, 32, 5, 6, 160, 3, 1), (2, 32, 9, 8, 64, 1, 2)]:
        x = ctx.to_device(rng.random((B, Cin, H, W), dtype=np.float32))
        w = ctx.to_device((rng.standard_normal((Cout, Cin, k, k)) / 10).astype(np.float32))
        b = ctx.to_device(np.zeros(Cout, np.float32))
        y = ctx.conv2d_forward(x, w, b, st)
        ctx.conv2d_backward(x, w, torch.randn_like(y), st)
    for (B, C, H, W) in [(2, 16, 111, 111), (1, 3, 8, 9), (2, 5, 31, 128)]:
        x = ctx.to_device(rng.standard_normal((B, C, H, W)).astype(np.float32))
        yr, yp, mask = ctx.relu_maxpool_forward(x, 2, 2)
        ctx.maxpool_relu_backward(torch.randn_like(yp), mask, yp, x.shape, 2, 2)
    init = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "alexnet_init.model"), np.float32)
    net = Net(ctx, alexnet_lite(3), 2)
    net.set_params(init)
    x, lab = ctx.to_device(synth_images(2)), ctx.to_device(synth_labels(2), torch.int32)
    for _ in range(3):   # eager warm-up step, graph capture, graph replay
        net.train_step(x, lab, 1e-3)
    ctx.sync()
    net.layer_output(0)          # lazy head: materialise on demand
    net.input_grad()
    net.set_lazy(False)
    for _ in range(2):
        net.train_step(x, lab, 1e-3)
    ctx.sync()
    print("memcheck workload done, loss", float(net.loss_from_slab()))
    net.close()
    from cnn_b200 import nets
    for spec, hw in ((nets.vgg_style(3, in_hw=76, width=32, hidden=32), 76), ):
        n2 = Net(ctx, spec, 2, 3, hw, hw)
        n2.set_params(nets.scaled_init(spec, seed=1))
        x2 = ctx.to_device(synth_images(2, 3, hw, hw))
        for _ in range(3):
            n2.train_step(x2, lab, 1e-3)
        ctx.sync()
        n2.input_grad()
        print("small VGG-style loss", float(n2.loss_from_slab()))
        n2.close()
    spec = nets.alexnet_lite(3, batch_norm=True)
    n3 = Net(ctx, spec, 2)
    n3.set_params(nets.insert_bn_params(spec, init))
    for _ in range(3):
        n3.train_step(x, lab, 1e-3)
    ctx.sync()
    print("AlexNet+BN loss", float(n3.loss_from_slab()))
    n3.close()
    ctx.close()


if __name__ == "__main__":
    main()
