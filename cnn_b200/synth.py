"""Portable synthetic inputs for benches and parity tests.

Pixels are i.i.d. U[0,1) drawn from a counter-based hash -- splitmix64(seed*2^40 + index),
top 24 bits / 2^24 -- so every value is exactly representable in fp32 and identical in
numpy, C++ and CUDA (cnn_b200/csrc/synth.cuh uses the same constants).  Range matches the
reference's u8/255 pixels (data_format.cpp:17-21).  Labels are b mod classes.
"""
import numpy as np

_M1 = np.uint64(0x9E3779B97F4A7C15)
_M2 = np.uint64(0xBF58476D1CE4E5B9)
_M3 = np.uint64(0x94D049BB133111EB)


def splitmix64(z):
    z = np.asarray(z, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = z + _M1
        z = (z ^ (z >> np.uint64(30))) * _M2
        z = (z ^ (z >> np.uint64(27))) * _M3
        return z ^ (z >> np.uint64(31))


def synth_uniform(n, seed, offset=0):
    """n fp32 values in [0,1); element i depends only on (seed, offset+i)."""
    idx = np.arange(offset, offset + n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = splitmix64((np.uint64(seed) << np.uint64(40)) + idx)
    return ((h >> np.uint64(40)).astype(np.float32)) * np.float32(1.0 / (1 << 24))


def synth_images(B, C=3, H=224, W=224, seed=1234, first_image=0):
    """[B,C,H,W] fp32; image b of a global batch is independent of how the batch is sharded."""
    per = C * H * W
    return synth_uniform(B * per, seed, first_image * per).reshape(B, C, H, W)


def synth_labels(B, classes=3, first_image=0):
    return ((np.arange(B, dtype=np.int64) + first_image) % classes).astype(np.int32)
