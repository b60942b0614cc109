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
from fp64_ref import fp64_step


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
    # fp64 tie-breaker: the reference's own sequential fp32 sums over 3 M terms are ~7e-5 from exact on the
    # first layer's weight gradient; a tensor passes within 1e-4 of the oracle or at least as close to fp64 as the oracle is
    _, _, g64 = fp64_step(spec, init, xh, lh)
    worst = {}
    for name, algo in (("default", api.CONV_AUTO), ("cuda-core", api.CONV_SIMT)):
        ctx.set_conv_algo(algo)
        net = Net(ctx, spec, B)
        net.set_params(init)
        net.train_step(x, lab, 1e-3)
        ctx.sync()
        g = net.get_grads()
        errs, bad = [], 0
        for li, kind, off, n in param_layout(spec)[0]:
            sl = slice(off, off + n)
            e = float(np.abs(g[sl] - g_ref[sl]).max() / np.abs(g_ref[sl]).max())
            e64 = float(np.abs(g[sl] - g64[sl]).max() / np.abs(g64[sl]).max())
            o64 = float(np.abs(g_ref[sl] - g64[sl]).max() / np.abs(g64[sl]).max())
            errs.append((f"L{li}.{kind}", e, e64, o64))
            bad += 0 if (e <= 1e-4 or e64 <= max(o64, 1e-4)) else 1
        worst[name] = (max(e for _, e, _, _ in errs), bad)
        print(f"{name:10s} loss {float(net.loss_from_slab()):.6f} grads vs oracle [vs fp64 | oracle vs fp64]: " +
              " ".join(f"{k} {e:.1e} [{a:.1e}|{b:.1e}]" for k, e, a, b in errs))
        net.close()
    print("FULLSTEP_PARITY", "OK" if worst["default"][1] == 0 else "FAILED", worst)


if __name__ == "__main__":
    main()
