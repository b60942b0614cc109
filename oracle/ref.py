"""ctypes binding of oracle/_ref/libcnn_ref.so -- the reference's OWN sources (compiled in
place from /root/reference/cpu/src by oracle/Makefile) behind oracle/ref_driver.cpp.
TEST INFRASTRUCTURE ONLY.  `available()` is False where the library was never built
(it cannot be rebuilt on the GPU box: /root/reference does not exist there; the prebuilt
.so travels with the snapshot).
"""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libcnn_ref.so")
_lib = None
_f = C.POINTER(C.c_float)
_i = C.POINTER(C.c_int)


def build(ref_root="/root/reference"):
    """Compile the reference sources in place (no-op message if they are absent)."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={ref_root}"])


def available():
    return os.path.exists(_LIB)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError("oracle/_ref/libcnn_ref.so not built (make -C oracle ref)")
        _lib = C.CDLL(_LIB)
        _lib.ref_alexnet_create.restype = C.c_void_p
        _lib.ref_net_create.restype = C.c_void_p
        _lib.ref_net_get_grads.restype = C.c_long
        _lib.ref_net_train_step.restype = C.c_float
        _lib.ref_net_layer_output.restype = C.c_long
    return _lib


def _fp(a):
    return None if a is None else a.ctypes.data_as(_f)


def _ip(a):
    return None if a is None else a.ctypes.data_as(_i)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def conv_out(h, k, s):
    return (h - k) // s + 1


def conv2d(x, w, bias, stride, delta=None):
    x, w, bias = _c(x), _c(w), _c(bias)
    B, Cin, H, W = x.shape
    Cout, _, k, _ = w.shape
    y = np.empty((B, Cout, conv_out(H, k, stride), conv_out(W, k, stride)), np.float32)
    if delta is None:
        lib().ref_conv2d(B, Cin, H, W, Cout, k, stride, _fp(x), _fp(w), _fp(bias), None,
                         _fp(y), None, None, None)
        return y
    delta = _c(delta)
    dw, db, dx = np.empty_like(w), np.empty(Cout, np.float32), np.empty_like(x)
    lib().ref_conv2d(B, Cin, H, W, Cout, k, stride, _fp(x), _fp(w), _fp(bias), _fp(delta),
                     _fp(y), _fp(dw), _fp(db), _fp(dx))
    return y, dw, db, dx


def maxpool(x, k, step, delta=None):
    x = _c(x)
    B, Cc, H, W = x.shape
    OH, OW = conv_out(H, k, step), conv_out(W, k, step)
    y = np.empty((B, Cc, OH, OW), np.float32)
    mask = np.empty((B, Cc, OH, OW), np.int32)
    dx = np.empty_like(x) if delta is not None else None
    d = _c(delta) if delta is not None else None
    lib().ref_maxpool(B, Cc, H, W, k, step, _fp(x), _fp(d), _fp(y), _ip(mask), _fp(dx))
    return y, mask, dx


def relu(x, delta=None):
    x = _c(x)
    B, Cc, H, W = x.shape
    y = np.empty_like(x)
    d = _c(delta).copy() if delta is not None else None
    lib().ref_relu(B, Cc, H, W, _fp(x), _fp(d), _fp(y))
    return y, d


def linear(x, w, bias, delta=None):
    x, w, bias = _c(x), _c(w), _c(bias)
    B, Cc, H, W = x.shape
    n_out = w.shape[1]
    y = np.empty((B, n_out), np.float32)
    if delta is None:
        lib().ref_linear(B, Cc, H, W, n_out, _fp(x), _fp(w), _fp(bias), None, _fp(y),
                         None, None, None)
        return y
    delta = _c(delta)
    dw, db, dx = np.empty_like(w), np.empty(n_out, np.float32), np.empty_like(x)
    lib().ref_linear(B, Cc, H, W, n_out, _fp(x), _fp(w), _fp(bias), _fp(delta), _fp(y),
                     _fp(dw), _fp(db), _fp(dx))
    return y, dw, db, dx


def batchnorm(x, gamma, beta, moving_mean, moving_var, delta=None, eval_mode=False):
    x, gamma, beta = _c(x), _c(gamma), _c(beta)
    mm, mv = _c(moving_mean).copy(), _c(moving_var).copy()
    B, Cc, H, W = x.shape
    y, xhat = np.empty_like(x), np.empty_like(x)
    mean, var = np.zeros(Cc, np.float32), np.zeros(Cc, np.float32)
    dg, dbt = np.zeros(Cc, np.float32), np.zeros(Cc, np.float32)
    d = _c(delta).copy() if delta is not None else None
    lib().ref_batchnorm(B, Cc, H, W, int(eval_mode), _fp(x), _fp(gamma), _fp(beta), _fp(mm),
                        _fp(mv), _fp(d), _fp(y), _fp(xhat), _fp(mean), _fp(var), _fp(dg), _fp(dbt))
    return dict(y=y, xhat=xhat, mean=mean, var=var, moving_mean=mm, moving_var=mv,
                dx=d, dgamma=dg, dbeta=dbt)


def softmax_xent(logits, labels=None):
    z = _c(logits)
    B, n = z.shape
    probs = np.empty_like(z)
    pred = np.empty(B, np.int32)
    if labels is None:
        lib().ref_softmax_xent(B, n, _fp(z), None, _fp(probs), None, None, _ip(pred))
        return probs, pred
    lab = np.ascontiguousarray(labels, np.int32)
    delta = np.empty_like(z)
    loss = C.c_float(0)
    lib().ref_softmax_xent(B, n, _fp(z), _ip(lab), _fp(probs), _fp(delta), C.byref(loss), _ip(pred))
    return probs, pred, np.float32(loss.value), delta


class Net:
    """Reference AlexNet (specs=None) or a list of reference layers built from specs."""

    def __init__(self, specs=None, num_classes=3, batch_norm=False):
        if specs is None:
            self._h = C.c_void_p(lib().ref_alexnet_create(num_classes, int(batch_norm)))
        else:
            flat = [int(v) for s in specs for v in (list(s) + [0] * 5)[:5]]
            arr = (C.c_int * len(flat))(*flat)
            self._h = C.c_void_p(lib().ref_net_create(arr, len(specs)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.ref_net_destroy(self._h)
            self._h = None

    def load_file(self, path):
        assert lib().ref_net_load_file(self._h, str(path).encode()) == 0

    def save_file(self, path):
        assert lib().ref_net_save_file(self._h, str(path).encode()) == 0

    def set_params(self, flat):
        with tempfile.NamedTemporaryFile(suffix=".model") as f:
            _c(flat).tofile(f.name)
            self.load_file(f.name)

    def get_params(self):
        with tempfile.NamedTemporaryFile(suffix=".model") as f:
            self.save_file(f.name)
            return np.fromfile(f.name, np.float32)

    def get_grads(self):
        n = lib().ref_net_get_grads(self._h, None)
        out = np.empty(n, np.float32)
        lib().ref_net_get_grads(self._h, _fp(out))
        return out

    def forward(self, x, classes=3, no_grad=False):
        x = _c(x)
        B, Cc, H, W = x.shape
        out = np.empty((B, classes), np.float32)
        n = lib().ref_net_forward(self._h, B, Cc, H, W, _fp(x), _fp(out), int(no_grad))
        assert n == classes, (n, classes)
        return out

    def train_step(self, x, labels, lr, classes=3, want_dx=False):
        x = _c(x)
        B, Cc, H, W = x.shape
        lab = np.ascontiguousarray(labels, np.int32)
        probs = np.empty((B, classes), np.float32)
        dx = np.empty_like(x) if want_dx else None
        loss = lib().ref_net_train_step(self._h, B, Cc, H, W, _fp(x), _ip(lab), C.c_float(lr),
                                        _fp(probs), _fp(dx))
        return np.float32(loss), probs, dx

    def grad_cam(self, x1, layer="conv_layer_3"):
        """AlexNet::grad_cam (alexnet.cpp:95-142) of one 3x224x224 image: (u8 map [H*W], probabilities)."""
        x1 = _c(x1.astype(np.float32))
        cam = np.zeros(1024, np.uint8)
        probs = np.zeros(3, np.float32)
        fn = lib().ref_alexnet_grad_cam
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_void_p, C.c_int, C.c_void_p]
        n = fn(self._h, x1.ctypes.data_as(C.c_void_p), layer.encode(), cam.ctypes.data_as(C.c_void_p), cam.size,
               probs.ctypes.data_as(C.c_void_p))
        assert n > 0, n
        return cam[:n].copy(), probs

    def layer_output(self, idx, B):
        n = lib().ref_net_layer_output(self._h, idx, B, None)
        out = np.empty(n, np.float32)
        lib().ref_net_layer_output(self._h, idx, B, _fp(out))
        return out


def alexnet_init_params(num_classes=3, batch_norm=False):
    """Initial parameters as the reference constructors draw them (seeds 212 / 1998)."""
    with tempfile.NamedTemporaryFile(suffix=".model") as f:
        lib().ref_alexnet_init_params_to_file(num_classes, int(batch_norm), f.name.encode())
        return np.fromfile(f.name, np.float32)
