#!/usr/bin/env python
"""Focused compute-sanitizer workload for the kernels added at the end of round 2: thin stride-1 first layer (forward +
weight gradient), two-warp stride-1 GEMM, one-launch Linear backward, zero padding / average pooling layers.
    compute-sanitizer --tool memcheck --error-exitcode 9 python tools/memcheck_new.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cnn_b200 import nets
from cnn_b200.api import Context, Net
from cnn_b200.synth import synth_images, synth_labels


def main():
    ctx = Context(0)
    rng = np.random.default_rng(0)
    for (B, H, W, Cout) in [(2, 9, 11, 16), (1, 7, 37, 64), (3, 5, 5, 32)]:
        x = ctx.to_device(rng.random((B, 3, H, W), dtype=np.float32))
        w = ctx.to_device((rng.standard_normal((Cout, 3, 3, 3)) / 10).astype(np.float32))
        b = ctx.to_device(np.zeros(Cout, np.float32))
        y = ctx.conv2d_forward(x, w, b, 1)
        ctx.conv2d_backward(x, w, torch.randn_like(y), 1)
    x = ctx.to_device(rng.random((2, 3, 6, 7), dtype=np.float32))
    yp = ctx.pad2d_forward(x, 2)
    ctx.pad2d_backward(yp, x.shape, 2)
    ya = ctx.avgpool_forward(x, 3, 2)
    ctx.avgpool_backward(ya, x.shape, 3, 2)
    lab = ctx.to_device(synth_labels(2), torch.int32)
    for spec, hw in ((nets.padded_resnet_shaped(3, width=32, in_hw=16), 16), (nets.vgg_style(3, in_hw=76, width=32, hidden=32), 76)):
        n = Net(ctx, spec, 2, 3, hw, hw)
        n.set_params(nets.scaled_init(spec, seed=1))
        xx = ctx.to_device(synth_images(2, 3, hw, hw))
        for _ in range(3):      # eager, capture, replay
            n.train_step(xx, lab, 1e-3)
        ctx.sync()
        print("loss", float(n.loss_from_slab()))
        n.close()
    ctx.close()
    print("memcheck_new workload done")


if __name__ == "__main__":
    main()
