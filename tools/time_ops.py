#!/usr/bin/env python
"""CUDA-event timing of the conv operators of one layer shape (experiments).
    python tools/time_ops.py conv0 [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as CT
import torch
from cnn_b200.api import Context, CONV_SIMT
from tools.prof_ops import SHAPES

def main():
    which = sys.argv[1]; B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    ctx = Context(0)
    if "simt" in sys.argv: ctx.set_conv_algo(CONV_SIMT)
    if "tc" in sys.argv: ctx.set_conv_algo(2)
    Cin, H, W, Cout, k, s = SHAPES[which]
    x = torch.rand(B, Cin, H, W, device="cuda"); w = torch.randn(Cout, Cin, k, k, device="cuda") / 10
    b = torch.zeros(Cout, device="cuda")
    y = ctx.conv2d_forward(x, w, b, s); d = torch.randn_like(y)
    dw, db, dx = torch.empty_like(w), torch.empty(Cout, device="cuda"), torch.empty_like(x)
    P = lambda t: CT.c_void_p(t.data_ptr()); L, h = ctx.L, ctx._h
    ops = {"fwd": lambda: L.cnn_conv2d_forward(h, P(x), P(w), P(b), P(y), B, Cin, H, W, Cout, k, s),
           "wgrad": lambda: L.cnn_conv2d_backward_weights(h, P(x), P(d), P(dw), P(db), B, Cin, H, W, Cout, k, s, 1.0 / B),
           "dgrad": lambda: L.cnn_conv2d_backward_data(h, P(w), P(d), P(dx), B, Cin, H, W, Cout, k, s)}
    out = []
    for name, fn in ops.items():
        with torch.cuda.stream(ctx.stream):
            fn(); fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(ctx.stream)
            for _ in range(5): fn()
            e1.record(ctx.stream)
        e1.synchronize()
        out.append(f"{name} {e0.elapsed_time(e1) / 5 * 1000:.1f}us")
    env = {k: v for k, v in os.environ.items() if k.startswith("CNN_DBG")}
    print(which, env, " ".join(out))

if __name__ == "__main__":
    main()
