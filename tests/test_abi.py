"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU and exports exactly
the entry points include/dnpy_b200.h declares (no compute is issued here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dnpy_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dnb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for required in ("dnb_dense_fwd", "dnb_dense_bwd", "dnb_conv2d_fwd", "dnb_conv2d_bwd", "dnb_dwconv2d_fwd",
                     "dnb_maxpool2d_fwd", "dnb_maxpool2d_bwd", "dnb_avgpool2d_fwd", "dnb_batchnorm_fwd",
                     "dnb_batchnorm_bwd", "dnb_softmax_fwd", "dnb_loss_ce", "dnb_loss_bce", "dnb_loss_mse",
                     "dnb_embedding_fwd", "dnb_embedding_bwd", "dnb_rnn_fwd", "dnb_rnn_bwd", "dnb_opt_adam",
                     "dnb_opt_rmsprop", "dnb_opt_sgd", "dnb_fill_f32"):
        assert required in syms


def test_library_exports_every_declared_symbol():
    from dnpy_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libdnpy_b200.so not built: run __graft_entry__.build()"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_python_binding_matches_header():
    from dnpy_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.dnb_version() >= 100
    assert isinstance(lib.dnb_last_error(), bytes)


def test_pure_host_entry_points():
    """Size queries are host-only arithmetic and safe without a device."""
    from dnpy_b200 import _lib
    lib = _lib.load()
    g = _lib.Win(8, 3, 7, 7, 3, 3, 3, 3, 3, 2, 2, 0, 0, 0, 0)
    assert lib.dnb_maxpool2d_bwd_ws_bytes(g, _lib.ROUTE_LITERAL) == 3 * 3 * 3 * (9 + 1) * 4   # last[npos][P] + done[npos]
    assert lib.dnb_maxpool2d_bwd_ws_bytes(g, _lib.ROUTE_PER_SAMPLE) == 0
    assert lib.dnb_loss_ws_bytes(10) == 14 * 4
    assert lib.dnb_embedding_bwd_ws_bytes(5, 7, 3) == (7 + 15 * (1 + 32)) * 4        # last[V] + mean[L*D] + 32 batch-slice partials
    assert lib.dnb_rnn_bwd_tc_ws_bytes(2, 3, 32) == 768                               # pre[B*T*H] fp32, rounded up to 256 B
    # the tensor-core variants are math-mode gated: fp32 never takes them (no device or driver needed to answer)
    assert lib.dnb_rnn_bwd_tc_supported(256, 8, 32, 4, _lib.MATH_FP32) == 0
    gc = _lib.Win(8, 1, 28, 28, 32, 26, 26, 3, 3, 1, 1, 0, 0, 0, 0)
    gp = _lib.Win(8, 32, 26, 26, 32, 12, 12, 3, 3, 2, 2, 0, 0, 0, 0)
    assert lib.dnb_conv_pool_act_tc_supported(gc, gp, _lib.MATH_FP32) == 0


def test_no_cpu_fallback_without_device():
    """Without a GPU a compute request must fail loudly, not silently run elsewhere."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    from tests import topologies
    ns = topologies.product_ns()
    l0 = ns.L.Input(shape=(4,))
    l0.output = np.ones((2, 4))
    l1 = ns.L.Relu(l0)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        l1.forward()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "dnpy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "/root/reference" not in src, f
