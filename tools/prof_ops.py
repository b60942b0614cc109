#!/usr/bin/env python
"""Runs single hot-path operators at BASELINE.json config-2 shapes (B=256) for ncu captures.

    ncu --set full -k regex:gather_gemm -c 3 python tools/prof_ops.py conv0
"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from cnn_b200.api import Context, CONV_SIMT

SHAPES = {  # name: (Cin, H, W, Cout, k, s)
    "conv0": (3, 224, 224, 16, 3, 2), "conv3": (16, 55, 55, 32, 3, 2),
    "conv5": (32, 27, 27, 64, 3, 2), "conv7": (64, 13, 13, 128, 3, 2),
}


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "conv0"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    ctx = Context(0)
    if "simt" in sys.argv:
        ctx.set_conv_algo(CONV_SIMT)
    Cin, H, W, Cout, k, s = SHAPES[which]
    x = torch.rand(B, Cin, H, W, device="cuda")
    w = torch.randn(Cout, Cin, k, k, device="cuda") / 10
    b = torch.zeros(Cout, device="cuda")
    for _ in range(reps):
        y = ctx.conv2d_forward(x, w, b, s)
        d = torch.randn_like(y)
        ctx.conv2d_backward(x, w, d, s)
    ctx.sync()
    print("ok", which, tuple(y.shape))


if __name__ == "__main__":
    main()
