#!/usr/bin/env python
"""Full-size (BASELINE.json config 3, B=128) parity of the tensor-core conv paths against the fp32
CUDA-core kernels of the same library (the CPU oracle would need hours at this size): normwise
max|a-b| / max|b| per tensor, bar 1e-4.    python tools/vgg_parity_fullsize.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from cnn_b200 import api
from cnn_b200.api import Context


def main():
    ctx = Context(0)
    torch.manual_seed(0)
    worst = 0.0
    shapes = [(128, 64, 222, 222, 64, 1), (128, 128, 108, 108, 128, 1), (128, 512, 22, 22, 512, 1)]
    if "--alexnet" in sys.argv:   # BASELINE.json config 2: the four AlexNet-lite layers at B=256
        shapes = [(256, 3, 224, 224, 16, 2), (256, 16, 55, 55, 32, 2), (256, 32, 27, 27, 64, 2), (256, 64, 13, 13, 128, 2)]
    for (B, Cin, H, W, Cout, st) in shapes:
        x = torch.rand(B, Cin, H, W, device="cuda")
        w = torch.randn(Cout, Cin, 3, 3, device="cuda") * (2.0 / (Cin * 9)) ** 0.5
        b = torch.zeros(Cout, device="cuda")
        out = {}
        for algo in ("simt", "auto"):
            ctx.set_conv_algo(api.CONV_SIMT if algo == "simt" else api.CONV_AUTO)
            y = ctx.conv2d_forward(x, w, b, st)
            if algo == "simt":
                d = torch.randn_like(y)
            dw, db, dx = ctx.conv2d_backward(x, w, d, st)
            ctx.sync()
            out[algo] = (y, dw, db, dx)
        errs = [float((a - r).abs().max() / r.abs().max()) for a, r in zip(out["auto"], out["simt"])]
        print(f"B={B} {Cin}->{Cout} {H}x{W} s{st}: tensor-core vs CUDA-core rel.err y {errs[0]:.2e} dw {errs[1]:.2e} db {errs[2]:.2e} dx {errs[3]:.2e}")
        # who is off?  fp64 evaluation of the weight / bias gradient (library cuDNN call, checker only)
        dw64 = torch.nn.grad.conv2d_weight(x.double(), w.shape, d.double(), stride=st) / B
        db64 = d.double().sum(dim=(0, 2, 3)) / B
        e = {k: (float((out[k][1].double() - dw64).abs().max() / dw64.abs().max()),
                 float((out[k][2].double() - db64).abs().max() / db64.abs().max())) for k in ("simt", "auto")}
        print(f"    against fp64: CUDA-core dw {e['simt'][0]:.2e} db {e['simt'][1]:.2e} | tensor-core dw {e['auto'][0]:.2e} db {e['auto'][1]:.2e}")
        worst = max(worst, errs[0], errs[3], e["auto"][0], e["auto"][1])
        del dw64, db64
        del out, x, y, d, dw, dx
        torch.cuda.empty_cache()
    print("VGG_FULLSIZE_PARITY", "OK" if worst <= 1e-4 else "FAILED", f"worst {worst:.2e}")


if __name__ == "__main__":
    main()
