"""Loader for the in-tree C-ABI library cnn_b200/libcnn_b200.so (include/cnn_b200.h).

There is no fallback of any kind: if the library is missing or a call fails, this raises.
"""
import ctypes as C
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(_HERE)
LIB_PATH = os.path.join(_HERE, "libcnn_b200.so")
HEADER = os.path.join(ROOT, "include", "cnn_b200.h")

_P, _I, _F, _Z, _LL = C.c_void_p, C.c_int, C.c_float, C.c_size_t, C.c_longlong

# name -> (restype, argtypes); mirrors include/cnn_b200.h one to one
SIGNATURES = {
    "cnn_last_error": (C.c_char_p, []),
    "cnn_version": (C.c_char_p, []),
    "cnn_ctx_create": (_I, [_I, _P, C.POINTER(_P)]),
    "cnn_ctx_destroy": (_I, [_P]),
    "cnn_ctx_set_stream": (_I, [_P, _P]),
    "cnn_ctx_stream": (_P, [_P]),
    "cnn_ctx_bind_numa": (_I, [_P]),
    "cnn_ctx_set_conv_algo": (_I, [_P, _I]),
    "cnn_ctx_set_tc_precision": (_I, [_P, _I]),
    "cnn_sync": (_I, [_P]),
    "cnn_launch_count": (_LL, [_P]),
    "cnn_prof_begin": (_I, [_P]),
    "cnn_prof_end": (_I, [_P, _P, _Z, _P, _I, C.POINTER(_I)]),
    "cnn_malloc": (_I, [_P, _Z, C.POINTER(_P)]),
    "cnn_free": (_I, [_P, _P]),
    "cnn_host_alloc": (_I, [_P, _Z, C.POINTER(_P)]),
    "cnn_host_free": (_I, [_P, _P]),
    "cnn_memset": (_I, [_P, _P, _I, _Z]),
    "cnn_h2d": (_I, [_P, _P, _P, _Z]),
    "cnn_d2h": (_I, [_P, _P, _P, _Z]),
    "cnn_d2d": (_I, [_P, _P, _P, _Z]),
    "cnn_conv2d_forward": (_I, [_P, _P, _P, _P, _P] + [_I] * 7),
    "cnn_conv2d_relu_maxpool_forward": (_I, [_P] * 8 + [_I] * 9),
    "cnn_conv2d_backward_weights": (_I, [_P, _P, _P, _P, _P] + [_I] * 7 + [_F]),
    "cnn_conv2d_backward_data": (_I, [_P, _P, _P, _P] + [_I] * 7),
    "cnn_maxpool_forward": (_I, [_P, _P, _P, _P] + [_I] * 6),
    "cnn_maxpool_backward": (_I, [_P, _P, _P, _P] + [_I] * 6),
    "cnn_relu_maxpool_forward": (_I, [_P, _P, _P, _P, _P] + [_I] * 6),
    "cnn_maxpool_relu_backward": (_I, [_P, _P, _P, _P, _P] + [_I] * 6),
    "cnn_relu_forward": (_I, [_P, _P, _P, _Z]),
    "cnn_relu_backward": (_I, [_P, _P, _P, _Z]),
    "cnn_linear_forward": (_I, [_P, _P, _P, _P, _P, _I, _I, _I]),
    "cnn_linear_backward": (_I, [_P] * 7 + [_I, _I, _I, _F]),
    "cnn_bn_forward_train": (_I, [_P] * 10 + [_I] * 4 + [_F, _F]),
    "cnn_bn_forward_eval": (_I, [_P] * 8 + [_I] * 4 + [_F]),
    "cnn_bn_backward": (_I, [_P] * 9 + [_I] * 4 + [_F]),
    "cnn_softmax_xent": (_I, [_P] * 7 + [_I, _I]),
    "cnn_xent_backward": (_I, [_P, _P, _P, _P, _P, _I, _I]),
    "cnn_sgd_step": (_I, [_P, _P, _P, _Z, _F]),
    "cnn_pad2d_forward": (_I, [_P, _P, _P] + [_I] * 5),
    "cnn_pad2d_backward": (_I, [_P, _P, _P] + [_I] * 5),
    "cnn_avgpool_forward": (_I, [_P, _P, _P] + [_I] * 6),
    "cnn_avgpool_backward": (_I, [_P, _P, _P] + [_I] * 6),
    "cnn_sgd_momentum_step": (_I, [_P, _P, _P, _P, _Z, _F, _F]),
    "cnn_adam_step": (_I, [_P, _P, _P, _P, _P, _Z, _F, _F, _F, _F, _I]),
    "cnn_net_create": (_I, [_P, C.POINTER(_I), _I, _I, _I, _I, _I, C.POINTER(_P)]),
    "cnn_net_destroy": (_I, [_P]),
    "cnn_net_param_count": (_LL, [_P]),
    "cnn_net_num_classes": (_I, [_P]),
    "cnn_net_params": (_P, [_P]),
    "cnn_net_grads": (_P, [_P]),
    "cnn_net_grad_slab_count": (_LL, [_P]),
    "cnn_net_set_params_host": (_I, [_P, _P]),
    "cnn_net_get_params_host": (_I, [_P, _P]),
    "cnn_net_get_grads_host": (_I, [_P, _P]),
    "cnn_net_use_graph": (_I, [_P, _I]),
    "cnn_net_forward": (_I, [_P, _P, _I]),
    "cnn_net_logits": (_P, [_P]),
    "cnn_net_probs": (_P, [_P]),
    "cnn_net_layer_output_host": (_I, [_P, _I, _P, C.POINTER(_LL)]),
    "cnn_net_backward": (_I, [_P, _P, _F]),
    "cnn_net_input_grad": (_P, [_P]),
    "cnn_net_update": (_I, [_P, _F]),
    "cnn_net_set_lazy": (_I, [_P, _I]),
    "cnn_net_enable_peer_exchange": (_I, [_P]),
    "cnn_net_materialize": (_I, [_P]),
    "cnn_net_pool_mask": (_P, [_P, _I]),
    "cnn_net_train_step": (_I, [_P, _P, _P, _F, _F, _I]),
    "cnn_net_train_step_host": (_I, [_P, _P, _P, _F, _P, _P]),
    "cnn_net_predict_host": (_I, [_P, _P, _P, _P]),
    "cnn_net_train_step_host_submit": (_I, [_P, _P, _P, _F]),
    "cnn_net_train_step_host_submit_u8": (_I, [_P, _P, _P, _F]),
    "cnn_net_train_step_host_wait": (_I, [_P, _P, _P]),
    "cnn_dist_unique_id": (_I, [_P]),
    "cnn_dist_init": (_I, [_P, _I, _I, _P]),
    "cnn_dist_world": (_I, [_P]),
    "cnn_dist_set_sync_bn": (_I, [_P, _I]),
    "cnn_dist_allreduce_sum": (_I, [_P, _P, _Z]),
    "cnn_dist_finalize": (_I, [_P]),
    "cnn_u8hwc_to_chw": (_I, [_P, _P, _P, _I, _I, _I, _I]),
}

_lib = None


def header_symbols():
    """Every function name include/cnn_b200.h declares."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"CNN_API[^;(]*?\b(cnn_\w+)\s*\(", f.read())))


def build(verbose=False):
    """nvcc cross-compile for sm_100a (works without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j8"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU/PyTorch fallback)")
        _lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
    return _lib


class CnnError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        raise CnnError(f"{what}: status {rc}: {lib().cnn_last_error().decode()}")
