"""Python face of the C ABI (include/cnn_b200.h) for tests and bench.py.

torch is used only for device memory and streams: every function here hands raw device
pointers of contiguous fp32 [B,C,H,W] torch tensors to libcnn_b200.so.  Operator names and
argument meaning follow the reference layer classes (cpu/include/architectures.h).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check

CONV_AUTO, CONV_SIMT, CONV_TCGEN05 = 0, 1, 2
TC_TF32X3, TC_BF16X3, TC_MIXED, TC_BF16X1 = 0, 1, 2, 3


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensor required"
    return C.c_void_p(t.data_ptr())


def _f32(t):
    assert t.dtype == torch.float32
    return _p(t)


def _i32(t):
    assert t is None or t.dtype == torch.int32
    return _p(t)


def conv_out(h, k, s):
    return (h - k) // s + 1


class Context:
    """cnn_ctx bound to a torch CUDA stream (a fresh side stream by default, so that the
    engine's CUDA-graph capture never touches the legacy default stream)."""

    def __init__(self, device=0, stream=None):
        if not torch.cuda.is_available():
            raise _lib.CnnError("cnn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = stream if stream is not None else torch.cuda.Stream(device=self.device)
        self._h = C.c_void_p()
        check(_lib.lib().cnn_ctx_create(device, C.c_void_p(self.stream.cuda_stream), C.byref(self._h)),
              "cnn_ctx_create")
        self.L = _lib.lib()
        self._nets = []  # weak refs: nets must be destroyed before their context

    def close(self):
        if self._h:
            for ref in self._nets:
                net = ref()
                if net is not None:
                    net.close()
            self._nets = []
            self.L.cnn_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        check(self.L.cnn_sync(self._h), "cnn_sync")

    def bind_numa(self):
        check(self.L.cnn_ctx_bind_numa(self._h), "cnn_ctx_bind_numa")

    def set_conv_algo(self, algo):
        check(self.L.cnn_ctx_set_conv_algo(self._h, algo), "cnn_ctx_set_conv_algo")

    def set_tc_precision(self, mode):
        check(self.L.cnn_ctx_set_tc_precision(self._h, mode), "cnn_ctx_set_tc_precision")

    @property
    def launches(self):
        return int(self.L.cnn_launch_count(self._h))

    def profile(self, fn):
        """Run fn() (eager launches on this context) and return [(kernel name, microseconds)] per launch."""
        check(self.L.cnn_prof_begin(self._h), "cnn_prof_begin")
        fn()
        names = C.create_string_buffer(1 << 16)
        us = (C.c_float * 512)()
        n = C.c_int(0)
        check(self.L.cnn_prof_end(self._h, names, len(names), us, 512, C.byref(n)), "cnn_prof_end")
        nm = names.value.decode().split("\n")
        return [(nm[i], float(us[i])) for i in range(min(n.value, len(nm)))]

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def to_device(self, a, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device).contiguous()

    # ---- operators (device tensors in, device tensors out) -------------------------
    def conv2d_forward(self, x, w, bias, stride):
        B, Cin, H, W = x.shape
        Cout, _, k, _ = w.shape
        y = self.empty(B, Cout, conv_out(H, k, stride), conv_out(W, k, stride))
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_conv2d_forward(self._h, _f32(x), _f32(w), _f32(bias), _f32(y), B, Cin, H, W,
                                            Cout, k, stride), "cnn_conv2d_forward")
        return y

    def conv2d_backward(self, x, w, delta, stride, scale=None):
        B, Cin, H, W = x.shape
        Cout, _, k, _ = w.shape
        scale = 1.0 / B if scale is None else scale
        dw, db, dx = torch.empty_like(w), self.empty(Cout), torch.empty_like(x)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_conv2d_backward_weights(self._h, _f32(x), _f32(delta), _f32(dw), _f32(db), B,
                                                     Cin, H, W, Cout, k, stride, scale),
                  "cnn_conv2d_backward_weights")
            check(self.L.cnn_conv2d_backward_data(self._h, _f32(w), _f32(delta), _f32(dx), B, Cin, H, W,
                                                  Cout, k, stride), "cnn_conv2d_backward_data")
        return dw, db, dx

    def maxpool_forward(self, x, k, step, want_mask=True):
        B, Cc, H, W = x.shape
        OH, OW = conv_out(H, k, step), conv_out(W, k, step)
        y = self.empty(B, Cc, OH, OW)
        mask = self.empty(B, Cc, OH, OW, dtype=torch.int32) if want_mask else None
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_maxpool_forward(self._h, _f32(x), _f32(y), _i32(mask), B, Cc, H, W, k, step),
                  "cnn_maxpool_forward")
        return y, mask

    def maxpool_backward(self, delta, mask, in_shape, k, step):
        B, Cc, H, W = in_shape
        dx = self.empty(B, Cc, H, W)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_maxpool_backward(self._h, _f32(delta), _i32(mask), _f32(dx), B, Cc, H, W, k,
                                              step), "cnn_maxpool_backward")
        return dx

    def relu_maxpool_forward(self, x, k, step):
        B, Cc, H, W = x.shape
        OH, OW = conv_out(H, k, step), conv_out(W, k, step)
        yr, yp = torch.empty_like(x), self.empty(B, Cc, OH, OW)
        mask = self.empty(B, Cc, OH, OW, dtype=torch.int32)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_relu_maxpool_forward(self._h, _f32(x), _f32(yr), _f32(yp), _i32(mask), B, Cc, H, W, k,
                                                  step), "cnn_relu_maxpool_forward")
        return yr, yp, mask

    def conv2d_relu_maxpool_forward(self, x, w, bias, stride, pool_k, pool_step):
        """Conv2D -> ReLU -> MaxPool2D in one kernel (the head of alexnet.cpp:12-16); returns
        (conv out, relu out, pool out, pool mask)."""
        B, Cin, H, W = x.shape
        Cout, _, k, _ = w.shape
        OH, OW = conv_out(H, k, stride), conv_out(W, k, stride)
        PH, PW = conv_out(OH, pool_k, pool_step), conv_out(OW, pool_k, pool_step)
        yc, yr = self.empty(B, Cout, OH, OW), self.empty(B, Cout, OH, OW)
        yp, mask = self.empty(B, Cout, PH, PW), self.empty(B, Cout, PH, PW, dtype=torch.int32)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_conv2d_relu_maxpool_forward(self._h, _f32(x), _f32(w), _f32(bias), _f32(yc), _f32(yr), _f32(yp),
                                                         _i32(mask), B, Cin, H, W, Cout, k, stride, pool_k, pool_step),
                  "cnn_conv2d_relu_maxpool_forward")
        return yc, yr, yp, mask

    def maxpool_relu_backward(self, delta, mask, pool_out, in_shape, k, step):
        B, Cc, H, W = in_shape
        dx = self.empty(B, Cc, H, W)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_maxpool_relu_backward(self._h, _f32(delta), _i32(mask), _f32(pool_out), _f32(dx), B, Cc,
                                                   H, W, k, step), "cnn_maxpool_relu_backward")
        return dx

    def relu_forward(self, x):
        y = torch.empty_like(x)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_relu_forward(self._h, _f32(x), _f32(y), x.numel()), "cnn_relu_forward")
        return y

    def relu_backward(self, delta, y):
        """In place on delta, like ReLU::backward (relu.cpp:30-44)."""
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_relu_backward(self._h, _f32(delta), _f32(y), delta.numel()), "cnn_relu_backward")
        return delta

    def linear_forward(self, x, w, bias):
        B = x.shape[0]
        n_in, n_out = w.shape
        y = self.empty(B, n_out)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_linear_forward(self._h, _f32(x), _f32(w), _f32(bias), _f32(y), B, n_in, n_out),
                  "cnn_linear_forward")
        return y

    def linear_backward(self, x, w, delta, scale=None):
        B = x.shape[0]
        n_in, n_out = w.shape
        scale = 1.0 / B if scale is None else scale
        dw, db, dx = torch.empty_like(w), self.empty(n_out), torch.empty_like(x)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_linear_backward(self._h, _f32(x), _f32(w), _f32(delta), _f32(dw), _f32(db),
                                             _f32(dx), B, n_in, n_out, scale), "cnn_linear_backward")
        return dw, db, dx

    def bn_forward_train(self, x, gamma, beta, moving_mean, moving_var, eps=1e-5, momentum=0.1):
        """moving_mean / moving_var are updated in place."""
        B, Cc, H, W = x.shape
        mean, var = self.empty(Cc), self.empty(Cc)
        xhat, y = torch.empty_like(x), torch.empty_like(x)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_bn_forward_train(self._h, _f32(x), _f32(gamma), _f32(beta), _f32(moving_mean),
                                              _f32(moving_var), _f32(mean), _f32(var), _f32(xhat), _f32(y),
                                              B, Cc, H, W, eps, momentum), "cnn_bn_forward_train")
        return dict(y=y, xhat=xhat, mean=mean, var=var)

    def bn_forward_eval(self, x, gamma, beta, moving_mean, moving_var, eps=1e-5):
        B, Cc, H, W = x.shape
        xhat, y = torch.empty_like(x), torch.empty_like(x)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_bn_forward_eval(self._h, _f32(x), _f32(gamma), _f32(beta), _f32(moving_mean),
                                             _f32(moving_var), _f32(xhat), _f32(y), B, Cc, H, W, eps),
                  "cnn_bn_forward_eval")
        return dict(y=y, xhat=xhat)

    def bn_backward(self, delta, x, xhat, gamma, mean, var, eps=1e-5):
        """In place on delta, like BatchNorm2D::backward."""
        B, Cc, H, W = x.shape
        dg, db = self.empty(Cc), self.empty(Cc)
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_bn_backward(self._h, _f32(delta), _f32(x), _f32(xhat), _f32(gamma), _f32(mean),
                                         _f32(var), _f32(dg), _f32(db), B, Cc, H, W, eps), "cnn_bn_backward")
        return delta, dg, db

    def softmax_xent(self, logits, labels=None):
        B, n = logits.shape
        probs = torch.empty_like(logits)
        pred = self.empty(B, dtype=torch.int32)
        delta = torch.empty_like(logits) if labels is not None else None
        loss_sum = self.empty(1) if labels is not None else None
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_softmax_xent(self._h, _f32(logits), _i32(labels), _f32(probs), _p(delta),
                                          _p(loss_sum), _i32(pred), B, n), "cnn_softmax_xent")
        return probs, pred, loss_sum, delta

    def sgd_step(self, params, grads, lr):
        with torch.cuda.stream(self.stream):
            check(self.L.cnn_sgd_step(self._h, _f32(params), _f32(grads), params.numel(), lr), "cnn_sgd_step")
        return params


def _opt_steps(Context):
    def sgd_momentum_step(self, params, grads, velocity, lr, momentum=0.9):
        check(self.L.cnn_sgd_momentum_step(self._h, _f32(params), _f32(grads), _f32(velocity), params.numel(), lr, momentum),
              "cnn_sgd_momentum_step")

    def adam_step(self, params, grads, m, v, lr, t, beta1=0.9, beta2=0.999, eps=1e-8):
        check(self.L.cnn_adam_step(self._h, _f32(params), _f32(grads), _f32(m), _f32(v), params.numel(), lr, beta1, beta2, eps, t),
              "cnn_adam_step")

    def avgpool_forward(self, x, k, step):
        B, Cc, H, W = x.shape
        y = self.empty(B, Cc, conv_out(H, k, step), conv_out(W, k, step))
        check(self.L.cnn_avgpool_forward(self._h, _f32(x), _f32(y), B, Cc, H, W, k, step), "cnn_avgpool_forward")
        return y

    def avgpool_backward(self, delta, in_shape, k, step):
        B, Cc, H, W = in_shape
        dx = self.empty(B, Cc, H, W)
        check(self.L.cnn_avgpool_backward(self._h, _f32(delta), _f32(dx), B, Cc, H, W, k, step), "cnn_avgpool_backward")
        return dx

    def pad2d_forward(self, x, pad):
        B, Cc, H, W = x.shape
        y = self.empty(B, Cc, H + 2 * pad, W + 2 * pad)
        check(self.L.cnn_pad2d_forward(self._h, _f32(x), _f32(y), B, Cc, H, W, pad), "cnn_pad2d_forward")
        return y

    def pad2d_backward(self, delta, in_shape, pad):
        B, Cc, H, W = in_shape
        dx = self.empty(B, Cc, H, W)
        check(self.L.cnn_pad2d_backward(self._h, _f32(delta), _f32(dx), B, Cc, H, W, pad), "cnn_pad2d_backward")
        return dx

    Context.pad2d_forward = pad2d_forward
    Context.pad2d_backward = pad2d_backward
    Context.avgpool_forward = avgpool_forward
    Context.avgpool_backward = avgpool_backward
    Context.sgd_momentum_step = sgd_momentum_step
    Context.adam_step = adam_step


_opt_steps(Context)


class Net:
    """cnn_net: the resident-buffer engine behind AlexNet::{forward,backward,update_gradients}."""

    def __init__(self, ctx, spec, B, Cc=3, H=224, W=224):
        self.ctx, self.B, self.in_shape = ctx, B, (B, Cc, H, W)
        self.L = ctx.L
        flat = [int(v) for s in spec for v in (list(s) + [0] * 5)[:5]]
        arr = (C.c_int * len(flat))(*flat)
        self._h = C.c_void_p()
        check(self.L.cnn_net_create(ctx._h, arr, len(spec), B, Cc, H, W, C.byref(self._h)), "cnn_net_create")
        self.n_params = int(self.L.cnn_net_param_count(self._h))
        self.classes = int(self.L.cnn_net_num_classes(self._h))
        self.n_layers = len(spec)
        import weakref
        ctx._nets.append(weakref.ref(self))

    def close(self):
        if getattr(self, "_h", None):
            if self.ctx._h:  # a net never outlives its context (Context.close destroys it first)
                self.L.cnn_net_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, ptr, count, dtype=torch.float32):
        """Zero-copy torch view of a library-owned device slab (plumbing for all-reduce)."""
        elt = 4

        class _Holder:
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f4" if dtype == torch.float32 else "<i4",
                                      "data": (int(ptr), False), "version": 2, "strides": (elt,)}
        return torch.as_tensor(h, device=self.ctx.device)

    def grad_slab(self):
        """[P+1] view: gradients in checkpoint order + the summed log-likelihood tail slot."""
        return self._view(self.L.cnn_net_grads(self._h), int(self.L.cnn_net_grad_slab_count(self._h)))

    def param_slab(self):
        return self._view(self.L.cnn_net_params(self._h), self.n_params)

    def probs(self):
        return self._view(self.L.cnn_net_probs(self._h), self.B * self.classes).view(self.B, self.classes)

    def logits(self):
        return self._view(self.L.cnn_net_logits(self._h), self.B * self.classes).view(self.B, self.classes)

    def input_grad(self):
        n = int(np.prod(self.in_shape))
        with torch.cuda.stream(self.ctx.stream):   # a lazy step materialises it here
            ptr = self.L.cnn_net_input_grad(self._h)
        assert ptr, _lib.lib().cnn_last_error().decode()
        return self._view(ptr, n).view(*self.in_shape)

    def set_params(self, flat):
        flat = np.ascontiguousarray(flat, np.float32)
        assert flat.size == self.n_params, (flat.size, self.n_params)
        check(self.L.cnn_net_set_params_host(self._h, flat.ctypes.data_as(C.c_void_p)), "set_params")

    def get_params(self):
        out = np.empty(self.n_params, np.float32)
        check(self.L.cnn_net_get_params_host(self._h, out.ctypes.data_as(C.c_void_p)), "get_params")
        return out

    def get_grads(self):
        out = np.empty(self.n_params, np.float32)
        check(self.L.cnn_net_get_grads_host(self._h, out.ctypes.data_as(C.c_void_p)), "get_grads")
        return out

    def use_graph(self, on):
        check(self.L.cnn_net_use_graph(self._h, int(on)), "use_graph")

    def enable_peer_exchange(self):
        """Collective over the ranks of init_native_dist: gradient sum + SGD as one peer-memory kernel.  False (on all
        ranks) where peer mapping is unavailable; the step then keeps ncclAllReduce + SGD."""
        rc = self.L.cnn_net_enable_peer_exchange(self._h)
        if rc == -4:   # CNN_ERR_UNSUPPORTED
            return False
        check(rc, "cnn_net_enable_peer_exchange")
        return True

    def set_lazy(self, on):
        """Lazy head of train steps (default on): head layer outputs / pool mask / image gradient on demand."""
        check(self.L.cnn_net_set_lazy(self._h, int(on)), "set_lazy")

    def materialize(self):
        with torch.cuda.stream(self.ctx.stream):
            check(self.L.cnn_net_materialize(self._h), "materialize")

    def pool_mask(self, idx, count):
        with torch.cuda.stream(self.ctx.stream):
            ptr = self.L.cnn_net_pool_mask(self._h, idx)
        assert ptr, "not a pool layer"
        return self._view(ptr, count, torch.int32)

    def forward(self, x, no_grad=False):
        with torch.cuda.stream(self.ctx.stream):
            check(self.L.cnn_net_forward(self._h, _f32(x), int(no_grad)), "cnn_net_forward")
        return self.logits()

    def backward(self, labels, grad_scale=None):
        scale = 1.0 / self.B if grad_scale is None else grad_scale
        with torch.cuda.stream(self.ctx.stream):
            check(self.L.cnn_net_backward(self._h, _i32(labels), scale), "cnn_net_backward")

    def update(self, lr):
        with torch.cuda.stream(self.ctx.stream):
            check(self.L.cnn_net_update(self._h, lr), "cnn_net_update")

    def train_step(self, x, labels, lr, grad_scale=None, do_update=True):
        scale = 1.0 / self.B if grad_scale is None else grad_scale
        with torch.cuda.stream(self.ctx.stream):
            check(self.L.cnn_net_train_step(self._h, _f32(x), _i32(labels), lr, scale, int(do_update)),
                  "cnn_net_train_step")

    def train_step_host(self, host_x, host_labels, lr, host_probs=None):
        """host_x: pinned torch/numpy fp32 [B,C,H,W]; returns the reference's loss value."""
        loss = C.c_float(0)
        check(self.L.cnn_net_train_step_host(self._h, _hp(host_x), _hp(host_labels), lr, C.byref(loss),
                                             _hp(host_probs)), "cnn_net_train_step_host")
        return np.float32(loss.value)

    def submit_host(self, host_x, host_labels, lr):
        """Pipelined host step: enqueue H2D (copy stream) + step; host buffers must stay alive until wait_host."""
        name = "cnn_net_train_step_host_submit_u8" if _is_u8(host_x) else "cnn_net_train_step_host_submit"
        check(getattr(self.L, name)(self._h, _hp(host_x), _hp(host_labels), lr), name)

    def wait_host(self, host_probs=None):
        loss = C.c_float(0)
        check(self.L.cnn_net_train_step_host_wait(self._h, C.byref(loss), _hp(host_probs)),
              "cnn_net_train_step_host_wait")
        return np.float32(loss.value)

    def predict_host(self, host_x):
        probs = np.empty((self.B, self.classes), np.float32)
        pred = np.empty(self.B, np.int32)
        check(self.L.cnn_net_predict_host(self._h, _hp(host_x), _hp(probs), _hp(pred)), "cnn_net_predict_host")
        return probs, pred

    def loss_from_slab(self, total_batch=None):
        """-sum/B as func.cpp:71 computes it (double arithmetic), from the grad-slab tail."""
        s = float(self.grad_slab()[-1].item())
        return np.float32(np.float64(np.float32(s)) * -1.0 / (total_batch or self.B))

    def layer_output(self, idx):
        cnt = C.c_longlong(0)
        self.ctx.sync()
        check(self.L.cnn_net_layer_output_host(self._h, idx, None, C.byref(cnt)), "layer_output")
        out = np.empty(cnt.value, np.float32)
        check(self.L.cnn_net_layer_output_host(self._h, idx, out.ctypes.data_as(C.c_void_p), C.byref(cnt)),
              "layer_output")
        return out


def _is_u8(a):
    return a.dtype == (torch.uint8 if isinstance(a, torch.Tensor) else np.uint8)


def _hp(a):
    """Host pointer of a numpy array or CPU torch tensor."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        assert not a.is_cuda and a.is_contiguous()
        return C.c_void_p(a.data_ptr())
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
