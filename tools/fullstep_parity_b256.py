#!/usr/bin/env python
"""Whole train step at BASELINE.json config 2 size (B=256) against the CPU oracle (~20 s of one host
core): per-tensor normwise error of every gradient for the default kernels and for the fp32 CUDA-core
convolution kernels.    python tools/fullstep_parity_b256.py [B]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cnn_b200 import api
from cnn_b200.api import Context, Net
from cnn_b200.nets import alexnet_lite, param_layout
from cnn_b200.synth import synth_images, synth_labels
from oracle import port


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    spec = alexnet_lite(3)
    init = np.fromfile(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "alexnet_init.model"), np.float32)
    xh, lh = synth_images(B, 3, 224, 224, seed=77), synth_labels(B, 3)
    t0 = time.time()
    o = port.Net(spec, B, 3, 224, 224)
    o.set_params(init)
    loss_ref, _, _ = o.train_step(xh, lh, 1e-3)
    g_ref = o.get_grads()
    print(f"oracle step at B={B}: {time.time() - t0:.1f} s, loss {loss_ref:.6f}")
    ctx = Context(0)
    x, lab = ctx.to_device(xh), ctx.to_device(lh, torch.int32)
    worst = {}
    for name, algo in (("default", api.CONV_AUTO), ("cuda-core", api.CONV_SIMT)):
        ctx.set_conv_algo(algo)
        net = Net(ctx, spec, B)
        net.set_params(init)
        net.train_step(x, lab, 1e-3)
        ctx.sync()
        g = net.get_grads()
        errs = []
        for li, kind, off, n in param_layout(spec)[0]:
            e = float(np.abs(g[off:off + n] - g_ref[off:off + n]).max() / np.abs(g_ref[off:off + n]).max())
            errs.append((f"L{li}.{kind}", e))
        worst[name] = max(e for _, e in errs)
        print(f"{name:10s} loss {float(net.loss_from_slab()):.6f} grads: " + " ".join(f"{k} {e:.1e}" for k, e in errs))
        net.close()
    print("FULLSTEP_PARITY", "OK" if worst["default"] <= 1e-4 else "FAILED", worst)


if __name__ == "__main__":
    main()
