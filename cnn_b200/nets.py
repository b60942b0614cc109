"""Layer-list descriptions of the networks on the hot path.

A spec is a list of (type, a, b, c, d) tuples with the layer codes of include/cnn_b200.h:
  CONV   (cin, cout, k, stride)   Conv2D ctor, architectures.h:69 (default k=3, stride=2, no padding)
  BN     (channels,)              BatchNorm2D, eps 1e-5, momentum 0.1 (architectures.h:167)
  RELU   ()
  POOL   (k, step)                MaxPool2D
  LINEAR (in, out)                LinearLayer
Extensions (items 7-8 of the reference's TODO list, cnn.cpp:15-24 -- not in the reference):
  PAD     (border,)               zero padding in front of a convolution
  AVGPOOL (k, step)               average pooling (k = input size: global pool)
Parameter order inside a flat slab is the reference checkpoint order (alexnet.cpp:69-77).
"""
import numpy as np

CONV, BN, RELU, POOL, LINEAR, PAD, AVGPOOL = 0, 1, 2, 3, 4, 5, 6


def alexnet_lite(num_classes=3, batch_norm=False):
    """The reference's hard-wired model, alexnet.cpp:12-31: 4x(conv k3 s2 [+BN] + ReLU), one
    MaxPool 2/2 after the first block, Linear 4608->classes.  Input 3x224x224."""
    spec = []
    for i, (cin, cout) in enumerate([(3, 16), (16, 32), (32, 64), (64, 128)]):
        spec.append((CONV, cin, cout, 3, 2))
        if batch_norm:
            spec.append((BN, cout, 0, 0, 0))
        spec.append((RELU, 0, 0, 0, 0))
        if i == 0:
            spec.append((POOL, 2, 2, 0, 0))
    spec.append((LINEAR, 6 * 6 * 128, num_classes, 0, 0))
    return spec


def vgg_style(num_classes=3, in_hw=224, width=64, hidden=256):
    """BASELINE.json config 3 / SURVEY §8d: eight 3x3 stride-1 convs (width*{1,1,2,2,4,4,8,8}
    channels), a 2/2 pool after every second conv, Linear(flat->hidden) ReLU
    Linear(hidden->classes); 224 -> 222 -> 220 -> 110 -> ... -> 10 at the default size."""
    spec = []
    mult = [1, 1, 2, 2, 4, 4, 8, 8]
    cin, hw = 3, in_hw
    for i, m in enumerate(mult):
        spec += [(CONV, cin, width * m, 3, 1), (RELU, 0, 0, 0, 0)]
        cin, hw = width * m, hw - 2
        if i % 2 == 1:
            spec.append((POOL, 2, 2, 0, 0))
            hw = (hw - 2) // 2 + 1
    spec += [(LINEAR, cin * hw * hw, hidden, 0, 0), (RELU, 0, 0, 0, 0), (LINEAR, hidden, num_classes, 0, 0)]
    return spec


def resnet18_shaped(num_classes=3):
    """BASELINE.json config 5 / SURVEY App. C: a ResNet-18-shaped stack the reference's layer API can express
    once 1x1 kernels are allowed (no residual adds, no padding): 3x3s2 3->64 (224->111), pool (->55),
    4x[3x3 64->64] (->47), 1x1s2 64->128 (->24), 4x[3x3 128->128] (->16), 1x1s2 128->256 (->8),
    2x[3x3 256->256] (->4), 1x1 256->512, 3x3 512->512 (->2), Linear 2048->classes; BN + ReLU after every conv."""
    spec = []

    def block(cin, cout, k, s):
        spec.extend([(CONV, cin, cout, k, s), (BN, cout, 0, 0, 0), (RELU, 0, 0, 0, 0)])

    block(3, 64, 3, 2)
    spec.append((POOL, 2, 2, 0, 0))
    for _ in range(4):
        block(64, 64, 3, 1)
    block(64, 128, 1, 2)
    for _ in range(4):
        block(128, 128, 3, 1)
    block(128, 256, 1, 2)
    for _ in range(2):
        block(256, 256, 3, 1)
    block(256, 512, 1, 1)
    block(512, 512, 3, 1)
    spec.append((LINEAR, 512 * 2 * 2, num_classes, 0, 0))
    return spec


def padded_resnet_shaped(num_classes=3, width=16, in_hw=32):
    """Small 'same'-padded net on the extension layers: PAD 1 + 3x3 stride-1 convolutions keep the resolution (what the
    reference's TODO item 8 asks for), BN + ReLU after each, max pool between stages, global average pool before the
    classifier (TODO item 7)."""
    spec, cin, hw = [], 3, in_hw
    for stage, cout in enumerate((width, 2 * width, 2 * width)):
        spec += [(PAD, 1, 0, 0, 0), (CONV, cin, cout, 3, 1), (BN, cout, 0, 0, 0), (RELU, 0, 0, 0, 0)]
        cin = cout
        if stage < 2:
            spec.append((POOL, 2, 2, 0, 0))
            hw //= 2
    spec += [(AVGPOOL, hw, hw, 0, 0), (LINEAR, cin, num_classes, 0, 0)]
    return spec


def scaled_init(spec, seed=0):
    """Fan-in scaled random parameters (N(0, sqrt(2/fan_in)) weights, zero biases, BN at its constructor state)
    for nets deeper than the reference's: its own N(0,1)/10 draws (conv2d.cpp:22-30) explode after a few wide layers
    (SURVEY 8d config 3)."""
    rng = np.random.default_rng(seed)
    lay, total = param_layout(spec)
    out = np.zeros(total, np.float32)
    for li, kind, off, n in lay:
        t, a, b, c, d = spec[li]
        if kind == "w":
            fan_in = a * c * c if t == CONV else a
            out[off:off + n] = rng.standard_normal(n).astype(np.float32) * np.float32(np.sqrt(2.0 / fan_in))
        elif kind == "gamma":
            out[off:off + n] = 1.0
    return out


def shapes(spec, C, H, W):
    """Per-layer (C,H,W) output shapes."""
    out = []
    for t, a, b, c, d in spec:
        if t == CONV:
            C, H, W = b, (H - c) // d + 1, (W - c) // d + 1
        elif t in (POOL, AVGPOOL):
            H, W = (H - a) // b + 1, (W - a) // b + 1
        elif t == PAD:
            H, W = H + 2 * a, W + 2 * a
        elif t == LINEAR:
            assert C * H * W == a, (C, H, W, a)
            C, H, W = b, 1, 1
        out.append((C, H, W))
    return out


def param_layout(spec):
    """[(layer_index, kind, offset, count)] in checkpoint order; kind in
    w,b (conv/linear), gamma,beta,moving_mean,moving_var (BN)."""
    lay, off = [], 0
    for i, (t, a, b, c, d) in enumerate(spec):
        if t == CONV:
            parts = [("w", b * a * c * c), ("b", b)]
        elif t == LINEAR:
            parts = [("w", a * b), ("b", b)]
        elif t == BN:
            parts = [("gamma", a), ("beta", a), ("moving_mean", a), ("moving_var", a)]
        else:
            parts = []
        for kind, n in parts:
            lay.append((i, kind, off, n))
            off += n
    return lay, off


def param_count(spec):
    return param_layout(spec)[1]


def insert_bn_params(spec_bn, flat_no_bn):
    """Takes a no-BN flat parameter vector and returns the flat vector for the same net with
    BatchNorm layers at their constructor state (gamma 1, beta 0, moving stats 0,
    batchnorm2d.cpp:17-21)."""
    lay, total = param_layout(spec_bn)
    out = np.zeros(total, np.float32)
    src = 0
    for _, kind, off, n in lay:
        if kind in ("w", "b"):
            out[off:off + n] = flat_no_bn[src:src + n]
            src += n
        elif kind == "gamma":
            out[off:off + n] = 1.0
    assert src == flat_no_bn.size
    return out


def flops_per_image(spec, C=3, H=224, W=224):
    """Forward MAC-FLOPs (2*MACs) per image of conv and linear layers; a train step is 3x."""
    tot = 0
    for (t, a, b, c, d), (oc, oh, ow) in zip(spec, shapes(spec, C, H, W)):
        if t == CONV:
            tot += 2 * oh * ow * b * a * c * c
        elif t == LINEAR:
            tot += 2 * a * b
    return tot
