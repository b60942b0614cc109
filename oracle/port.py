"""numpy/ctypes binding of the plain-C oracle (oracle/cnn_oracle.c).  TEST INFRASTRUCTURE ONLY.

All arrays are contiguous fp32 [B, C, H, W] (image b == the reference's b-th Tensor3D).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libcnn_oracle.so")

CONV, BN, RELU, POOL, LINEAR = 0, 1, 2, 3, 4


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "port"])


def _load():
    src = os.path.join(_HERE, "cnn_oracle.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        build()
    return C.CDLL(_LIB)


_lib = _load()
_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_i)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


_lib.orc_cross_entropy_backward.restype = C.c_float
_lib.orc_net_create.restype = C.c_void_p
_lib.orc_net_param_count.restype = C.c_long
_lib.orc_net_train_step.restype = C.c_float
_lib.orc_net_layer_output.restype = _f


def conv_out(h, k, s):
    return (h - k) // s + 1


def conv2d_forward(x, w, bias, stride):
    x, w, bias = _c(x), _c(w), _c(bias)
    B, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    y = np.empty((B, Cout, conv_out(H, k, stride), conv_out(W, k, stride)), np.float32)
    _lib.orc_conv2d_forward(_fp(x), _fp(w), _fp(bias), _fp(y), B, Cin, H, W, Cout, k, stride)
    return y


def conv2d_backward(x, w, delta, stride):
    x, w, delta = _c(x), _c(w), _c(delta)
    B, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    dw = np.empty_like(w)
    db = np.empty(Cout, np.float32)
    dx = np.empty_like(x)
    _lib.orc_conv2d_backward(_fp(x), _fp(w), _fp(delta), _fp(dw), _fp(db), _fp(dx),
                             B, Cin, H, W, Cout, k, stride)
    return dw, db, dx


def maxpool_forward(x, k, step, want_mask=True):
    x = _c(x)
    B, Cc, H, W = x.shape
    OH, OW = conv_out(H, k, step), conv_out(W, k, step)
    y = np.empty((B, Cc, OH, OW), np.float32)
    mask = np.zeros((B, Cc, OH, OW), np.int32) if want_mask else None
    _lib.orc_maxpool_forward(_fp(x), _fp(y), _ip(mask), B, Cc, H, W, k, step)
    return y, mask


def maxpool_backward(delta, mask, in_shape):
    delta = _c(delta)
    mask = np.ascontiguousarray(mask, np.int32)
    B, Cc, H, W = in_shape
    dx = np.empty(in_shape, np.float32)
    _lib.orc_maxpool_backward(_fp(delta), _ip(mask), _fp(dx), B, Cc, H, W,
                              delta.shape[2], delta.shape[3])
    return dx


# ---- extensions that are NOT in the reference (cnn.cpp:15-24 TODO items 7-8); see cnn_oracle.h
def pad_forward(x, pad):
    x = _c(x)
    B, Cc, H, W = x.shape
    y = np.empty((B, Cc, H + 2 * pad, W + 2 * pad), np.float32)
    _lib.orc_pad_forward(_fp(x), _fp(y), B, Cc, H, W, pad)
    return y


def pad_backward(delta, in_shape, pad):
    delta = _c(delta)
    B, Cc, H, W = in_shape
    dx = np.empty(in_shape, np.float32)
    _lib.orc_pad_backward(_fp(delta), _fp(dx), B, Cc, H, W, pad)
    return dx


def avgpool_forward(x, k, step):
    x = _c(x)
    B, Cc, H, W = x.shape
    y = np.empty((B, Cc, conv_out(H, k, step), conv_out(W, k, step)), np.float32)
    _lib.orc_avgpool_forward(_fp(x), _fp(y), B, Cc, H, W, k, step)
    return y


def avgpool_backward(delta, in_shape, k, step):
    delta = _c(delta)
    B, Cc, H, W = in_shape
    dx = np.empty(in_shape, np.float32)
    _lib.orc_avgpool_backward(_fp(delta), _fp(dx), B, Cc, H, W, k, step)
    return dx


def relu_forward(x):
    x = _c(x)
    y = np.empty_like(x)
    _lib.orc_relu_forward(_fp(x), _fp(y), C.c_long(x.size))
    return y


def relu_backward(delta, y):
    d = _c(delta).copy()
    y = _c(y)
    _lib.orc_relu_backward(_fp(d), _fp(y), C.c_long(d.size))
    return d


def linear_forward(x, w, bias):
    x, w, bias = _c(x), _c(w), _c(bias)
    B = x.shape[0]
    n_in, n_out = w.shape
    y = np.empty((B, n_out), np.float32)
    _lib.orc_linear_forward(_fp(x), _fp(w), _fp(bias), _fp(y), B, n_in, n_out)
    return y


def linear_backward(x, w, delta):
    x, w, delta = _c(x), _c(w), _c(delta)
    B = x.shape[0]
    n_in, n_out = w.shape
    dw = np.empty_like(w)
    db = np.empty(n_out, np.float32)
    dx = np.empty_like(x)
    _lib.orc_linear_backward(_fp(x), _fp(w), _fp(delta), _fp(dw), _fp(db), _fp(dx),
                             B, n_in, n_out)
    return dw, db, dx


def bn_forward_train(x, gamma, beta, moving_mean, moving_var, eps=1e-5, momentum=0.1):
    x, gamma, beta = _c(x), _c(gamma), _c(beta)
    mm, mv = _c(moving_mean).copy(), _c(moving_var).copy()
    B, Cc, H, W = x.shape
    mean = np.empty(Cc, np.float32)
    var = np.empty(Cc, np.float32)
    xhat = np.empty_like(x)
    y = np.empty_like(x)
    _lib.orc_bn_forward_train(_fp(x), _fp(gamma), _fp(beta), _fp(mm), _fp(mv), _fp(mean),
                              _fp(var), _fp(xhat), _fp(y), B, Cc, H, W,
                              C.c_float(eps), C.c_float(momentum))
    return dict(y=y, xhat=xhat, mean=mean, var=var, moving_mean=mm, moving_var=mv)


def bn_forward_eval(x, gamma, beta, moving_mean, moving_var, eps=1e-5):
    x, gamma, beta = _c(x), _c(gamma), _c(beta)
    mm, mv = _c(moving_mean), _c(moving_var)
    B, Cc, H, W = x.shape
    xhat = np.empty_like(x)
    y = np.empty_like(x)
    _lib.orc_bn_forward_eval(_fp(x), _fp(gamma), _fp(beta), _fp(mm), _fp(mv), _fp(xhat),
                             _fp(y), B, Cc, H, W, C.c_float(eps))
    return dict(y=y, xhat=xhat)


def bn_backward(delta, x, xhat, gamma, mean, var, eps=1e-5):
    d = _c(delta).copy()
    x, xhat, gamma, mean, var = _c(x), _c(xhat), _c(gamma), _c(mean), _c(var)
    B, Cc, H, W = x.shape
    dgamma = np.empty(Cc, np.float32)
    dbeta = np.empty(Cc, np.float32)
    _lib.orc_bn_backward(_fp(d), _fp(x), _fp(xhat), _fp(gamma), _fp(mean), _fp(var),
                         _fp(dgamma), _fp(dbeta), B, Cc, H, W, C.c_float(eps))
    return d, dgamma, dbeta


def softmax(logits):
    z = _c(logits)
    p = np.empty_like(z)
    _lib.orc_softmax(_fp(z), _fp(p), z.shape[0], z.shape[1])
    return p


def argmax_rows(v):
    v = _c(v)
    return np.array([_lib.orc_argmax(_fp(v[b]), v.shape[1]) for b in range(v.shape[0])], np.int32)


def cross_entropy_backward(probs, labels):
    p = _c(probs)
    lab = np.ascontiguousarray(labels, np.int32)
    d = np.empty_like(p)
    loss = _lib.orc_cross_entropy_backward(_fp(p), _ip(lab), _fp(d), p.shape[0], p.shape[1])
    return np.float32(loss), d


def sgd(p, g, lr):
    p = _c(p).copy()
    g = _c(g)
    _lib.orc_sgd(_fp(p), _fp(g), C.c_long(p.size), C.c_float(lr))
    return p


class Net:
    """orc_net wrapper.  specs: list of (type, a, b, c, d) tuples."""

    def __init__(self, specs, B, Cc, H, W):
        arr = (C.c_int * (5 * len(specs)))(*[int(v) for s in specs for v in (list(s) + [0] * 5)[:5]])
        self.B, self.shape = B, (B, Cc, H, W)
        self.n_layers = len(specs)
        self._h = C.c_void_p(_lib.orc_net_create(arr, len(specs), B, Cc, H, W))
        self.n_params = _lib.orc_net_param_count(self._h)
        self.classes = _lib.orc_net_num_classes(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            _lib.orc_net_destroy(self._h)
            self._h = None

    def set_params(self, flat):
        flat = _c(flat)
        assert flat.size == self.n_params, (flat.size, self.n_params)
        _lib.orc_net_set_params(self._h, _fp(flat))

    def get_params(self):
        out = np.empty(self.n_params, np.float32)
        _lib.orc_net_get_params(self._h, _fp(out))
        return out

    def get_grads(self):
        out = np.empty(self.n_params, np.float32)
        _lib.orc_net_get_grads(self._h, _fp(out))
        return out

    def forward(self, x, no_grad=False):
        x = _c(x)
        logits = np.empty((self.B, self.classes), np.float32)
        _lib.orc_net_forward(self._h, _fp(x), _fp(logits), int(no_grad))
        return logits

    def train_step(self, x, labels, lr, want_dx=False):
        x = _c(x)
        lab = np.ascontiguousarray(labels, np.int32)
        probs = np.empty((self.B, self.classes), np.float32)
        dx = np.empty(self.shape, np.float32) if want_dx else None
        loss = _lib.orc_net_train_step(self._h, _fp(x), _ip(lab), C.c_float(lr), _fp(probs), _fp(dx))
        return np.float32(loss), probs, dx

    def layer_output(self, idx):
        cnt = C.c_long(0)
        p = _lib.orc_net_layer_output(self._h, idx, C.byref(cnt))
        return np.ctypeslib.as_array(p, shape=(cnt.value,)).copy()
