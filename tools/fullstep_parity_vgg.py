#!/usr/bin/env python
"""BASELINE.json config 3 (VGG-style, eight 3x3 stride-1 convs, 224x224) at FULL image size: one whole train step
through the default (tensor-core) dispatch against the same step through the library's fp32 CUDA-core kernels
(CNN_CONV_SIMT) -- loss, probabilities and EVERY gradient / updated-parameter tensor within the 1e-4 bar,
normwise max|a-b| / max|b| per tensor.  The CPU oracle needs ~150 s per image at this size, so it checks
the same step for one image (--oracle-images N, default 1) of the batch: logits through the whole stack.
    python tools/fullstep_parity_vgg.py [--batch 16] [--oracle-images 1]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from cnn_b200 import api, nets
from cnn_b200.api import Context, Net
from cnn_b200.synth import synth_images, synth_labels
from fp64_ref import fp64_step


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--oracle-images", type=int, default=0)
    ap.add_argument("--no-fp64", action="store_true")
    a = ap.parse_args()
    B = a.batch
    spec = nets.vgg_style(3)
    params = nets.scaled_init(spec, seed=0)
    x, lab = synth_images(B, seed=11), synth_labels(B, 3)
    ctx = Context(0)
    res = {}
    for algo in ("simt", "auto"):
        ctx.set_conv_algo(api.CONV_SIMT if algo == "simt" else api.CONV_AUTO)
        net = Net(ctx, spec, B)
        net.use_graph(False)
        net.set_params(params)
        net.train_step(ctx.to_device(x), ctx.to_device(lab, torch.int32), 1e-3)
        ctx.sync()
        res[algo] = dict(loss=float(net.loss_from_slab()), probs=net.probs().cpu().numpy().copy(), logits=net.logits().cpu().numpy().copy(),
                         grads=net.get_grads(), params=net.get_params())
        net.close()
    s, t = res["simt"], res["auto"]
    worst = max(abs(t["loss"] - s["loss"]) / max(1.0, abs(s["loss"])), rel(t["probs"], s["probs"]))
    print(f"B={B}: loss {t['loss']:.6f} vs {s['loss']:.6f}; probs {rel(t['probs'], s['probs']):.2e} logits {rel(t['logits'], s['logits']):.2e}")
    g64 = None
    if not a.no_fp64:
        t0 = time.time()
        l64, lo64, g64 = fp64_step(spec, params, x, lab)
        print(f"fp64 evaluation ({time.time() - t0:.1f} s): loss {l64:.6f}; logits tensor-core {rel(t['logits'], lo64):.2e} CUDA-core {rel(s['logits'], lo64):.2e}")
    # Bar per tensor: 1e-4 against the fp32 CUDA-core path, or -- the fp64 tie-breaker of SURVEY 8(c) -- at least as close
    # to the fp64 evaluation as that fp32 path is.  (At this size the batch-mean gradients cancel by two to three orders
    # of magnitude: two correct fp32 evaluations that sum in a different order differ by ~1e-3.)
    bad = 0
    for li, kind, off, n in nets.param_layout(spec)[0]:
        eg, ep = rel(t["grads"][off:off + n], s["grads"][off:off + n]), rel(t["params"][off:off + n], s["params"][off:off + n])
        msg = f"  layer {li:2d} {kind}: grad {eg:.2e} param {ep:.2e}"
        ok = eg <= 1e-4 and ep <= 1e-4
        if g64 is not None:
            et, es = rel(t["grads"][off:off + n], g64[off:off + n]), rel(s["grads"][off:off + n], g64[off:off + n])
            msg += f" | grad vs fp64: tensor-core {et:.2e} CUDA-core {es:.2e}"
            ok = ok or et <= max(es, 1e-4)
            worst = max(worst, min(eg, et))
        else:
            worst = max(worst, eg)
        bad += 0 if ok else 1
        print(msg + ("" if ok else "   <-- outside the bar"))
    if a.oracle_images > 0:
        from oracle import port
        n = a.oracle_images
        t0 = time.time()
        o = port.Net(spec, n, 3, 224, 224)
        o.set_params(params)
        lo = o.forward(x[:n])
        e = rel(t["logits"][:n], lo)
        worst = max(worst, e)
        print(f"  CPU oracle forward of {n} image(s) ({time.time() - t0:.0f} s): logits {e:.2e}")
    print("VGG_FULLSTEP_PARITY", "OK" if bad == 0 else "FAILED", f"tensors outside the bar: {bad}; worst min(err vs CUDA-core, err vs fp64) {worst:.2e}")
    ctx.close()
    return 0 if bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
