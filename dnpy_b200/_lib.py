"""ctypes binding of ``libdnpy_b200.so`` (the C ABI declared in include/dnpy_b200.h).

The library is the ONLY compute path: there is no NumPy/torch fallback.  If the
shared object is missing, or a kernel is requested without a CUDA device, the
call fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdnpy_b200.so")

OK, ERR_INVALID, ERR_CUDA, ERR_UNSUPPORTED = 0, -1, -2, -3
MATH_FP32, MATH_TF32, MATH_BF16 = 0, 1, 2
ROUTE_LITERAL, ROUTE_PER_SAMPLE = 0, 1


class Reg(C.Structure):
    _fields_ = [("l1", C.c_float), ("l2", C.c_float)]


class Win(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "C", "H", "W", "F", "OH", "OW", "kh", "kw", "sy", "sx",
                                       "pt", "pb", "pl", "pr")]


_p, _i, _f, _z = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_W, _R = C.POINTER(Win), Reg

# name -> (restype, argtypes); mirrors include/dnpy_b200.h one to one
SIGNATURES = {
    "dnb_version": (_i, []),
    "dnb_last_error": (C.c_char_p, []),
    "dnb_device_query": (_i, [C.POINTER(_i)] * 4),
    "dnb_launch_count": (C.c_longlong, []),
    "dnb_prof_begin": (_i, [_p]),
    "dnb_prof_end": (_i, [C.c_char_p, _z]),
    "dnb_fill_f32": (_i, [_p, _z, _f, _p]),
    "dnb_fill_i32": (_i, [_p, _z, C.c_int32, _p]),
    "dnb_dense_fwd": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "dnb_dense_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _f, _R, _R, _f, _i, _p]),
    "dnb_dense_bwd_data": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "dnb_conv2d_bwd_data": (_i, [_p, _p, _p, _W, _p, _i, _p]),
    "dnb_conv2d_ws_bytes": (_z, [_W, _i]),
    "dnb_conv2d_fwd": (_i, [_p, _p, _p, _p, _W, _p, _i, _p]),
    "dnb_conv2d_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _W, _f, _R, _R, _f, _p, _i, _p]),
    "dnb_conv_pool_act_supported": (_i, [_W, _W]),
    "dnb_conv_pool_act_fwd": (_i, [_p, _p, _p, _p, _p, _W, _W, _f, _p]),
    "dnb_conv_pool_act_bwd": (_i, [_p, _p, _p, _p, _p, _W, _W, _f, _f, _p]),
    "dnb_conv_pool_act_bwd_literal_ws_bytes": (_z, [_W, _W]),
    "dnb_conv_pool_act_bwd_literal": (_i, [_p, _p, _p, _p, _p, _W, _W, _f, _f, _p, _p]),
    "dnb_conv_pool_act_tc_supported": (_i, [_W, _W, _i]),
    "dnb_conv_pool_act_fwd_tc": (_i, [_p, _p, _p, _p, _p, _W, _W, _f, _i, _p]),
    "dnb_conv2d_pool_act_supported": (_i, [_W, _W, _i]),
    "dnb_conv2d_pool_act_fwd": (_i, [_p, _p, _p, _p, _p, _W, _W, _f, _p, _i, _p]),
    "dnb_pool_act_supported": (_i, [_W]),
    "dnb_maxpool2d_act_fwd": (_i, [_p, _p, _p, _W, _f, _p]),
    "dnb_pool_act_bwd": (_i, [_p, _p, _p, _W, _f, _p]),
    "dnb_idx_gate_unpack": (_i, [_p, _p, _z, _p]),
    "dnb_idx_gate_set": (_i, [_p, _p, _z, _p]),
    "dnb_dwconv2d_fwd": (_i, [_p, _p, _p, _p, _W, _p]),
    "dnb_dwconv2d_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _W, _f, _R, _f, _p]),
    "dnb_maxpool2d_fwd": (_i, [_p, _p, _p, _W, _p]),
    "dnb_maxpool2d_bwd_ws_bytes": (_z, [_W, _i]),
    "dnb_maxpool2d_bwd": (_i, [_p, _p, _p, _W, _i, _p, _p]),
    "dnb_maxpool2d_bwd_prepare": (_i, [_p, _W, _p, _p]),
    "dnb_maxpool2d_bwd_apply": (_i, [_p, _p, _W, _p, _p]),
    "dnb_avgpool2d_fwd": (_i, [_p, _p, _W, _p]),
    "dnb_avgpool2d_bwd": (_i, [_p, _p, _W, _p]),
    "dnb_leakyrelu_fwd": (_i, [_p, _p, _z, _f, _p]),
    "dnb_leakyrelu_bwd": (_i, [_p, _p, _p, _z, _f, _p]),
    "dnb_prelu_fwd": (_i, [_p, _p, _p, _i, _z, _p]),
    "dnb_prelu_bwd": (_i, [_p, _p, _p, _p, _p, _i, _z, _f, _p]),
    "dnb_sigmoid_fwd": (_i, [_p, _p, _z, _p]),
    "dnb_sigmoid_bwd": (_i, [_p, _p, _p, _z, _p]),
    "dnb_tanh_fwd": (_i, [_p, _p, _z, _p]),
    "dnb_tanh_bwd": (_i, [_p, _p, _p, _z, _p]),
    "dnb_softmax_fwd": (_i, [_p, _p, _i, _i, _i, _p]),
    "dnb_softmax_bwd": (_i, [_p, _p, _p, _i, _i, _p]),
    "dnb_logsoftmax_fwd": (_i, [_p, _p, _i, _i, _i, _p]),
    "dnb_mul": (_i, [_p, _p, _p, _z, _p]),
    "dnb_scale": (_i, [_p, _p, _z, _f, _p]),
    "dnb_add_n": (_i, [C.POINTER(_p), _i, _p, _z, _p]),
    "dnb_merge_n": (_i, [C.POINTER(_p), _i, _p, _z, _i, _f, _p]),
    "dnb_batchnorm_fwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _z, _f, _f, _i, _p]),
    "dnb_batchnorm_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _i, _z, _p]),
    "dnb_batchnorm_stats": (_i, [_p, _p, _p, _i, _z, _p]),
    "dnb_batchnorm_fwd_synced": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _i, _z, _f, _f, _p]),
    "dnb_batchnorm_bwd_sums": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _z, _p]),
    "dnb_batchnorm_bwd_synced": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _z, _p]),
    "dnb_embedding_fwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "dnb_embedding_bwd_ws_bytes": (_z, [_i, _i, _i]),
    "dnb_embedding_bwd": (_i, [_p, _p, _p, _i, _i, _i, _i, _f, _p, _p]),
    "dnb_embedding_bwd_prepare": (_i, [_p, _p, _i, _i, _i, _i, _f, _i, _p, _p]),
    "dnb_embedding_bwd_apply": (_i, [_p, _p, _i, _i, _i, _f, _p]),
    "dnb_rnn_fwd": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "dnb_rnn_bwd": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "dnb_rnn_bwd_tc_supported": (_i, [_i, _i, _i, _i, _i]),
    "dnb_rnn_bwd_tc_ws_bytes": (_z, [_i, _i, _i]),
    "dnb_rnn_bwd_tc": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _p, _i, _p]),
    "dnb_loss_ws_bytes": (_z, [_i]),
    "dnb_loss_ce": (_i, [_p, _p, _p, _p, _i, _i, _i, _f, _p, _p]),
    "dnb_loss_bce": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "dnb_loss_mse": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "dnb_loss_rmse": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "dnb_loss_mae": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "dnb_loss_nll": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "dnb_loss_hinge": (_i, [_p, _p, _p, _p, _i, _i, _f, _p, _p]),
    "dnb_metric_categorical": (_i, [_p, _p, _p, _i, _i, _p]),
    "dnb_metric_binary": (_i, [_p, _p, _p, _i, _i, _f, _p]),
    "dnb_opt_adam": (_i, [_p, _p, _p, _p, _z, _f, _f, _f, _f, _p]),
    "dnb_opt_rmsprop": (_i, [_p, _p, _p, _z, _f, _f, _f, _p]),
    "dnb_opt_sgd": (_i, [_p, _p, _p, _z, _f, _f, _i, _p]),
    "dnb_dp_flag_bytes": (_z, []),
    "dnb_ipc_alloc": (_i, [C.POINTER(_p), _z, _p]),
    "dnb_ipc_open": (_i, [_p, C.POINTER(_p)]),
    "dnb_ipc_close": (_i, [_p]),
    "dnb_ipc_free": (_i, [_p]),
    "dnb_dp_allreduce": (_i, [C.POINTER(_p), C.POINTER(_p), _p, _z, _i, _i, _p]),
    "dnb_gemm_tf32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _p]),
}

_lib = None


def load():
    """Load the shared object (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"dnpy_b200: CUDA extension not built ({LIB_PATH} missing). Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` or dnpy_b200/csrc/build.sh — "
            "there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError = header/library drift: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    """Map a DNB_ERR_* code to the exception type the reference raises."""
    if rc == OK:
        return
    msg = load().dnb_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(f"dnpy_b200 CUDA error: {msg}")


def call(name, *args):
    check(getattr(load(), name)(*args))
