"""GPU parity tests proper (-m gpu): every §8a row through the dnpy-API layer classes -> ctypes ->
C ABI -> CUDA kernels, compared with (a) the fixtures recorded from the real reference and (b) the
oracle on larger seeded inputs.

Tolerances (FP32 path, math="fp32"): rel 1e-5 in the tensor-level metric of conftest.rel_err
(max|a-b| / max|b|); integer work (pool argmax, embedding gather, dropout masks) is bit-exact.
"""
import numpy as np
import pytest

from tests.conftest import rel_err, has_cuda
from tests.golden import cases, drive

pytestmark = pytest.mark.gpu

TOL32 = 1e-5


@pytest.fixture(autouse=True)
def _fp32_mode():
    import dnpy_b200
    dnpy_b200.set_math("fp32")
    dnpy_b200.set_maxpool_route("literal")
    yield


def test_native_library_is_the_compute_path():
    import torch
    from dnpy_b200 import _lib
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    lib = _lib.load()
    import ctypes as C
    sm, maj, mnr, tc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    _lib.check(lib.dnb_device_query(C.byref(sm), C.byref(maj), C.byref(mnr), C.byref(tc)))
    assert sm.value > 0 and maj.value >= 9
    loaded = open("/proc/self/maps").read()
    assert "libdnpy_b200.so" in loaded


@pytest.mark.parametrize("case", cases.LAYER_CASES, ids=[c["name"] for c in cases.LAYER_CASES])
def test_layer_cases_vs_reference_fixtures(case, golden, ns):
    g = golden("layer_cases.npz")
    rec = drive.run_layer_case(ns, case)
    n = case["name"]
    for k in [k.split("/", 1)[1] for k in g.files if k.startswith(n + "/")]:
        want = g[f"{n}/{k}"]
        if k.startswith(("x", "pin_", "delta")):
            continue
        got = np.asarray(rec[k])
        if k == "idx":
            assert np.array_equal(got, want), "argmax indices must be bit-exact"
        elif case["kind"] == "Embedding" and k == "out":
            assert np.array_equal(got.astype(np.float32), want.astype(np.float32)), "gather must be bit-exact"
        elif case["kind"] == "Dropout" and k == "out":
            assert np.array_equal(got == 0, want == 0), "dropout mask must be bit-exact"
            assert rel_err(got, want) <= TOL32
        else:
            assert rel_err(got.reshape(want.shape), want) <= TOL32, (k, rel_err(got.reshape(want.shape), want))


@pytest.mark.parametrize("case", cases.LOSS_CASES, ids=[c["name"] for c in cases.LOSS_CASES])
def test_loss_cases(case, golden, ns):
    g = golden("loss_cases.npz")
    rec = drive.run_loss_case(ns, case)
    n = case["name"]
    assert abs(float(rec["loss"]) - float(g[f"{n}/loss"])) <= TOL32 * max(1.0, abs(float(g[f"{n}/loss"])))
    assert rel_err(rec["delta"], g[f"{n}/delta"]) <= TOL32


@pytest.mark.parametrize("case", cases.OPT_CASES, ids=[c["name"] for c in cases.OPT_CASES])
def test_opt_cases(case, golden, ns):
    g = golden("opt_cases.npz")
    rec = drive.run_opt_case(ns, case)
    n = case["name"]
    for k in ("w1", "b1", "frozen_stat"):
        assert rel_err(rec[f"p_{k}"], g[f"{n}/p_{k}"]) <= TOL32, k


def test_reference_unit_kats_exact(golden, ns):
    """The reference's own unit tests (all-ones kernels, integer images): exact ``==`` like the
    reference asserts (tests/coverage/test_layers.py:78,82)."""
    import json
    L = ns.L
    g = golden("ref_unit_kat.npz")
    for i in range(int(g["n"])):
        cfg = json.loads(str(g[f"{i}_cfg"]))
        kind = cfg["kind"]
        if kind == "Softmax":
            continue
        l0 = L.Input(shape=tuple(cfg["in_shape"]))
        l0.output = g[f"{i}_x"]
        if kind == "Conv2D":
            l1 = L.Conv2D(l0, cfg["filters"], kernel_size=tuple(cfg["kernel_size"][-2:]), strides=tuple(cfg["strides"]),
                          padding=cfg["padding"], kernel_initializer=ns.init.Constant(fill_value=1))
        elif kind == "DepthwiseConv2D":
            l1 = L.DepthwiseConv2D(l0, kernel_size=tuple(cfg["kernel_size"][-2:]), strides=tuple(cfg["strides"]),
                                   padding=cfg["padding"], kernel_initializer=ns.init.Constant(fill_value=1))
        else:
            l1 = getattr(L, kind)(l0, pool_size=tuple(cfg["pool_size"]), strides=tuple(cfg["strides"]), padding=cfg["padding"])
        l1.initialize()
        l1.forward()
        assert np.all(g[f"{i}_out"] == l1.output), (i, kind)
        l1.delta = g[f"{i}_delta"]
        l1.backward()
        assert np.all(g[f"{i}_dx"] == l0.delta), (i, kind)


def _oracle_pair(ns, case, route="literal"):
    import dnpy_b200
    dnpy_b200.set_maxpool_route(route)
    got = drive.run_layer_case(ns, case)
    want = drive.run_layer_case_oracle(ns, case, maxpool_route=route)
    return got, want


BIG_CASES = [
    dict(name="big_dense", kind="Dense", in_shape=(784,), batch=256, kw=dict(units=512)),
    dict(name="big_dense_head", kind="Dense", in_shape=(1600,), batch=128, kw=dict(units=10, kernel_regularizer=["L2", 0.01])),
    # classifier-head kernels (out <= 16, in % 4 == 0): row tails, a W-chunk tail (520 = 512 + 8), out = 1 / 3 / 16
    dict(name="head_out3_rowtail", kind="Dense", in_shape=(300,), batch=70, kw=dict(units=3, kernel_regularizer=["L1", 0.02])),
    dict(name="head_out16_chunktail", kind="Dense", in_shape=(520,), batch=65, kw=dict(units=16)),
    dict(name="head_out1_wide", kind="Dense", in_shape=(4096,), batch=300, kw=dict(units=1)),
    dict(name="big_conv1", kind="Conv2D", in_shape=(1, 28, 28), batch=64,
         kw=dict(filters=32, kernel_size=(3, 3), strides=(1, 1), padding="none")),
    # flat thin-conv kernels (C <= 2, plane % 4 == 0): filter-count tails, padding, many images per CTA, W % 4 != 0
    dict(name="flat_conv_c2_same", kind="Conv2D", in_shape=(2, 12, 12), batch=37,
         kw=dict(filters=7, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="flat_conv_c1_same_many", kind="Conv2D", in_shape=(1, 8, 8), batch=700,
         kw=dict(filters=5, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="flat_conv_c1_w30", kind="Conv2D", in_shape=(1, 30, 30), batch=9,
         kw=dict(filters=12, kernel_size=(3, 3), strides=(1, 1), padding="none")),
    dict(name="big_conv2", kind="Conv2D", in_shape=(32, 12, 12), batch=48,
         kw=dict(filters=64, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="big_conv2_tall", kind="Conv2D", in_shape=(32, 28, 12), batch=5,       # several bands per image (no bulk path)
         kw=dict(filters=32, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="big_conv2_c64", kind="Conv2D", in_shape=(64, 12, 12), batch=7,        # two 32-channel blocks forward
         kw=dict(filters=16, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="big_conv_s2", kind="Conv2D", in_shape=(32, 16, 16), batch=16,
         kw=dict(filters=64, kernel_size=(3, 3), strides=(2, 2), padding="same")),
    dict(name="big_conv_c3", kind="Conv2D", in_shape=(3, 32, 32), batch=16,
         kw=dict(filters=32, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="big_pointwise", kind="PointwiseConv2D", in_shape=(32, 16, 16), batch=16, kw=dict(filters=32)),
    dict(name="ring_depthwise_nonsquare", kind="DepthwiseConv2D", in_shape=(3, 9, 12), batch=5,      # ring kernels: odd H, row tails
         kw=dict(kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="ring_depthwise_w8_many", kind="DepthwiseConv2D", in_shape=(5, 8, 8), batch=70,         # several planes per CTA
         kw=dict(kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="big_depthwise", kind="DepthwiseConv2D", in_shape=(32, 16, 16), batch=16,
         kw=dict(kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="big_maxpool", kind="MaxPool2D", in_shape=(32, 26, 26), batch=130,
         kw=dict(pool_size=(3, 3), strides=(2, 2), padding="none")),
    dict(name="big_maxpool_ties", kind="MaxPool2D", in_shape=(8, 12, 12), batch=200, quantize=True,
         kw=dict(pool_size=(3, 3), strides=(2, 2), padding="none")),
    # bulk-copy ring forward (unpadded, plane % 4 == 0): 2x2 pools, 5x5 outputs (chunk size from divisibility),
    # stride 1, a one-chunk launch; odd planes fall back to the strip / generic kernels
    dict(name="ring_maxpool_2x2", kind="MaxPool2D", in_shape=(16, 32, 32), batch=8,
         kw=dict(pool_size=(2, 2), strides=(2, 2), padding="none")),
    dict(name="ring_maxpool_5x5out", kind="MaxPool2D", in_shape=(64, 12, 12), batch=24, quantize=True,
         kw=dict(pool_size=(3, 3), strides=(2, 2), padding="none")),
    dict(name="ring_maxpool_s1", kind="MaxPool2D", in_shape=(4, 10, 10), batch=6,
         kw=dict(pool_size=(3, 3), strides=(1, 1), padding="none")),
    dict(name="ring_maxpool_many_chunks", kind="MaxPool2D", in_shape=(32, 12, 12), batch=1300,
         kw=dict(pool_size=(3, 3), strides=(2, 2), padding="none")),
    dict(name="maxpool_odd_plane", kind="MaxPool2D", in_shape=(3, 13, 13), batch=5,
         kw=dict(pool_size=(3, 3), strides=(2, 2), padding="none")),
    dict(name="big_avgpool", kind="AvgPool2D", in_shape=(16, 16, 16), batch=32,
         kw=dict(pool_size=(2, 2), strides=(2, 2), padding="none")),
    dict(name="big_batchnorm", kind="BatchNorm", in_shape=(8, 9, 9), batch=257, kw=dict(), training=True),
    dict(name="vec_batchnorm", kind="BatchNorm", in_shape=(16, 8, 12), batch=37, kw=dict(), training=True),     # 128-bit column path, odd row tail
    dict(name="big_relu", kind="Relu", in_shape=(32, 12, 12), batch=67, kw=dict()),
    dict(name="big_softmax", kind="Softmax", in_shape=(10,), batch=1000, kw=dict()),
    dict(name="big_softmax_wide", kind="Softmax", in_shape=(100,), batch=33, kw=dict()),
    dict(name="big_embedding", kind="Embedding", in_shape=(64,), batch=32, int_input=100,
         kw=dict(input_dim=100, output_dim=8, input_length=64)),
    # pscale 0.1: with 0.5*randn recurrent weights (spectral radius ~2.8) the recurrence is chaotic and
    # amplifies fp32-vs-fp64 round-off exponentially in T — a property of the test weights, not the kernel
    dict(name="big_rnn", kind="SimpleRNN", in_shape=(40, 8), batch=19, pscale=0.1, kw=dict(hidden_dim=32, bptt_truncate=4)),
    dict(name="warp_rnn_odd_dims", kind="SimpleRNN", in_shape=(7, 3), batch=11, pscale=0.1, kw=dict(hidden_dim=7, bptt_truncate=2)),
    dict(name="big_rnn_full_bptt", kind="SimpleRNN", in_shape=(12, 5), batch=9, pscale=0.1, kw=dict(hidden_dim=16)),
    dict(name="big_dropout", kind="Dropout", in_shape=(1000,), batch=16, kw=dict(rate=0.25), training=True, seed=3),
]


@pytest.mark.parametrize("case", BIG_CASES, ids=[c["name"] for c in BIG_CASES])
def test_config_shapes_vs_oracle(case, ns):
    got, want = _oracle_pair(ns, case)
    for k, w in want.items():
        if k not in got:
            continue
        gk = np.asarray(got[k])
        if k == "idx":
            assert np.array_equal(gk, w)
        else:
            e = rel_err(gk.reshape(np.shape(w)), w)
            assert e <= TOL32, (k, e)


# Persistent tensor-core kernels (grid = min(units, #SMs)): MANY units per CTA, so the multi-unit loop — smem ring reuse,
# mbarrier phase wrap, TMEM accumulator double-buffer reuse, `it >= 1` — is compared with the oracle, not only it == 0.
# B=600 -> ~4 images per CTA, B=2048 -> ~14; the 64x64 maps split into several row bands per image (cfg4's shapes).
PERSISTENT_CASES = [
    dict(name="sv_conv2_b600", kind="Conv2D", in_shape=(32, 12, 12), batch=600,
         kw=dict(filters=64, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="sv_conv2_b2048", kind="Conv2D", in_shape=(32, 12, 12), batch=2048,
         kw=dict(filters=64, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="sv_conv_s1_bands_b24", kind="Conv2D", in_shape=(32, 64, 64), batch=24,
         kw=dict(filters=32, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="sv_conv_s2_bands_b40", kind="Conv2D", in_shape=(32, 64, 64), batch=40,
         kw=dict(filters=64, kernel_size=(3, 3), strides=(2, 2), padding="same")),
    dict(name="sv_pointwise_b40", kind="PointwiseConv2D", in_shape=(32, 64, 64), batch=40, kw=dict(filters=32)),
    dict(name="tc_dense_b4096", kind="Dense", in_shape=(784,), batch=4096, kw=dict(units=512)),
    # SimpleRNN backward on tcgen05 (csrc/rnn_tc.cu): tiles of 124 owned rows cross sample boundaries at every phase; 362 / 166
    # tiles (several per CTA: operand buffers, mbarrier phases and the TMEM accumulator are reused); full-length lags (T <= 8)
    dict(name="tc_rnn_b700_t64", kind="SimpleRNN", in_shape=(64, 8), batch=700, pscale=0.1, kw=dict(hidden_dim=32, bptt_truncate=4)),
    dict(name="tc_rnn_b80_t256", kind="SimpleRNN", in_shape=(256, 8), batch=80, pscale=0.1, kw=dict(hidden_dim=32, bptt_truncate=4)),
    dict(name="tc_rnn_b1024_t256", kind="SimpleRNN", in_shape=(256, 8), batch=1024, pscale=0.1, kw=dict(hidden_dim=32, bptt_truncate=4)),
    dict(name="tc_rnn_b33_t7_full", kind="SimpleRNN", in_shape=(7, 24), batch=33, pscale=0.1, kw=dict(hidden_dim=32)),
    dict(name="tc_rnn_b5_t50_trunc1", kind="SimpleRNN", in_shape=(50, 3), batch=5, pscale=0.1, kw=dict(hidden_dim=32, bptt_truncate=1)),
]

TC_KINDS = ("Dense", "Conv2D", "PointwiseConv2D", "SimpleRNN")
TOL_TF32 = 2e-2


@pytest.mark.parametrize("math", ["tf32", "bf16"])
@pytest.mark.parametrize("case", [c for c in cases.LAYER_CASES + BIG_CASES + PERSISTENT_CASES if c["kind"] in TC_KINDS],
                         ids=lambda c: c["name"])
def test_tensor_core_path_vs_oracle(case, math, ns):
    """math="tf32": tcgen05 kernels (TF32 operands, fp32 accumulation in TMEM) for Dense / Conv2D /
    PointwiseConv2D, tolerance rel 2e-2 (BASELINE north_star); shapes the tensor path does not take
    (tiny reductions, N < 32) transparently run the fp32 kernels.  math="bf16": the same with BF16 operands in
    the shifted-view 3x3 convolution kernels (forward, data gradient, weight gradient)."""
    import dnpy_b200
    dnpy_b200.set_math(math)
    try:
        got = drive.run_layer_case(ns, case)
    finally:
        dnpy_b200.set_math("fp32")
    want = drive.run_layer_case_oracle(ns, case)
    for k, w in want.items():
        if k in got:
            e = rel_err(np.asarray(got[k]).reshape(np.shape(w)), w)
            assert e <= TOL_TF32, (k, e)


_FULL = {}


@pytest.mark.parametrize("math", ["tf32", "bf16"])
def test_full_batch_conv2_vs_oracle(math, ns):
    """The benchmarked launch itself: Conv2D 32x12x12 -> 64, 3x3 'same' at the BASELINE batch (B=8192: 55 images per
    persistent CTA) on the tcgen05 shifted-view kernels, forward / data gradient / weight + bias gradient against the
    float64 oracle on the SAME 8192 images (tolerance rel 2e-2, tensor-level metric of conftest.rel_err); a random
    sample of images is additionally checked image by image so an error confined to a few units cannot hide."""
    import dnpy_b200
    case = dict(name="sv_conv2_b8192", kind="Conv2D", in_shape=(32, 12, 12), batch=8192,
                kw=dict(filters=64, kernel_size=(3, 3), strides=(1, 1), padding="same"))
    if "want" not in _FULL:
        _FULL["want"] = drive.run_layer_case_oracle(ns, case)
    want = _FULL["want"]
    dnpy_b200.set_math(math)
    try:
        got = drive.run_layer_case(ns, case)
    finally:
        dnpy_b200.set_math("fp32")
    for k in ("out", "dx0", "g_w1", "g_b1"):
        e = rel_err(np.asarray(got[k]).reshape(np.shape(want[k])), want[k])
        assert e <= TOL_TF32, (k, e)
    pick = np.random.RandomState(0).choice(8192, 96, replace=False)
    for k in ("out", "dx0"):
        for b in pick:
            e = rel_err(got[k][b], want[k][b])
            assert e <= TOL_TF32, (k, int(b), e)


def test_tc_gemm_all_operand_orientations():
    """dnb_gemm_tf32: K-major and MN-major (128B swizzle with 32-byte atoms) operands, edge tiles."""
    import torch
    from dnpy_b200 import _lib
    lib = _lib.load()
    torch.manual_seed(0)
    for (M, N, K) in [(128, 128, 32), (256, 32, 96), (1024, 512, 784), (200, 100, 52), (1000, 36, 40)]:
        A, B = torch.randn(M, K, device="cuda"), torch.randn(K, N, device="cuda")
        ref = A.double() @ B.double()
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                a_st = A.t().contiguous() if a_mn else A.contiguous()
                b_st = B.contiguous() if b_mn else B.t().contiguous()
                C = torch.full((M, N), float("nan"), device="cuda")
                _lib.check(lib.dnb_gemm_tf32(a_st.data_ptr(), b_st.data_ptr(), C.data_ptr(), M, N, K, a_mn, b_mn, None))
                torch.cuda.synchronize()
                err = float((C.double() - ref).abs().max() / ref.abs().max())
                assert err <= 5e-3, (M, N, K, a_mn, b_mn, err)


@pytest.mark.parametrize("B,F,alpha", [(37, 32, 0.0), (4001, 32, 0.1), (1500, 64, 0.0), (700, 128, 0.3)])
def test_tensor_core_first_block_forward(B, F, alpha):
    """dnb_conv_pool_act_fwd_tc (csrc/tc_conv_c1p.cu): Conv2D(1 -> F, 3x3) -> MaxPool2D(3x3 / 2) -> LeakyRelu on tcgen05 with
    BF16 operands.  BF16 x BF16 products are exact in fp32, so against the float64 oracle evaluated on the BF16-ROUNDED image
    and filters (the bias is carried exactly) only the fp32 accumulation order is left: values to 1e-5, argmax bytes equal
    except at near-ties, the gate bit consistent with the value.  B = 4001: 1001 groups of four images (every CTA loops
    over ~7 groups: raw / tile / accumulator rings wrap, both epilogue groups) with a partial last group; F = 64 / 128:
    two / one image per MMA.  Against the un-rounded oracle (what the net-level tests see): rel 2e-2."""
    import torch
    from dnpy_b200 import _lib
    from oracle import dnpy_oracle as O
    lib = _lib.load()
    rs = np.random.RandomState(B + F)
    x = rs.rand(B, 1, 28, 28).astype(np.float32)
    w = (rs.randn(F, 1, 3, 3) * 0.4).astype(np.float32)
    b = (rs.randn(F) * 0.1).astype(np.float32)
    g = _lib.Win(B, 1, 28, 28, F, 26, 26, 3, 3, 1, 1, 0, 0, 0, 0)
    gp = _lib.Win(B, F, 26, 26, F, 12, 12, 3, 3, 2, 2, 0, 0, 0, 0)
    assert lib.dnb_conv_pool_act_tc_supported(g, gp, _lib.MATH_BF16) == 1
    assert lib.dnb_conv_pool_act_tc_supported(g, gp, _lib.MATH_FP32) == 0
    xd, wd, bd = (torch.from_numpy(a).cuda() for a in (x, w, b))
    act = torch.full((B, F, 12, 12), float("nan"), device="cuda")
    ig = torch.full((B, F, 12, 12), 255, dtype=torch.uint8, device="cuda")
    _lib.check(lib.dnb_conv_pool_act_fwd_tc(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), act.data_ptr(), ig.data_ptr(),
                                            g, gp, alpha, _lib.MATH_BF16, None))
    torch.cuda.synchronize()
    act, ig = act.cpu().numpy(), ig.cpu().numpy()

    def oracle(xx, ww):
        conv = O.conv2d_forward(xx.astype(np.float64), ww.astype(np.float64), b.astype(np.float64), (1, 1), (0, 0, 0, 0), (F, 26, 26))
        pooled, idx = O.maxpool2d_forward(conv, (3, 3), (2, 2), (0, 0, 0, 0), (F, 12, 12))
        return O.leakyrelu_forward(pooled, alpha), idx, pooled
    rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    want, idx, pooled = oracle(rnd(x), rnd(w))
    assert rel_err(act, want) <= 1e-5
    assert float(np.mean((ig & 15) != idx)) <= 2e-4
    clear = np.abs(pooled) > 1e-4
    assert np.array_equal(((ig >> 4) & 1)[clear], (pooled > 0)[clear].astype(np.uint8))
    assert np.all((ig >> 5) == 0)
    want32, idx32, _ = oracle(x, w)
    assert rel_err(act, want32) <= TOL_TF32
    assert float(np.mean((ig & 15) != idx32)) <= 0.03


@pytest.mark.parametrize("case", [c for c in cases.LAYER_CASES + BIG_CASES if c["kind"] in ("MaxPool2D", "GlobalMaxPool2D")],
                         ids=lambda c: c["name"])
def test_maxpool_per_sample_route(case, ns):
    got, want = _oracle_pair(ns, case, route="per_sample")
    assert np.array_equal(np.asarray(got["idx"]), want["idx"])
    assert rel_err(got["dx0"], want["dx0"]) <= TOL32


def test_empty_batch_is_a_noop(ns):
    L = ns.L
    l0 = L.Input(shape=(5,))
    l0.output = np.zeros((0, 5))
    l1 = L.Relu(l0)
    l1.forward()
    assert np.asarray(l1.output).shape == (0, 5)


def test_bad_arguments_raise_reference_exception_types(ns):
    from dnpy_b200 import _lib
    from dnpy_b200.darray import dev_empty, stream_ptr
    buf = dev_empty((16,))
    with pytest.raises(ValueError):
        _lib.call("dnb_loss_bce", buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), buf.data_ptr(), 4, 2, 0.25,
                  buf.data_ptr(), stream_ptr())
    g = _lib.Win(1, 1, 6, 6, 1, 2, 2, 3, 3, 2, 2, 0, 0, 0, 0)
    with pytest.raises(ValueError):      # the reference's AvgPool backward also raises for 3x3 pools
        _lib.call("dnb_avgpool2d_bwd", buf.data_ptr(), buf.data_ptr(), g, stream_ptr())


def test_nan_guard(ns):
    lo = ns.losses.CrossEntropy()
    net = ns.Net()
    p = np.array([[0.5, 0.5], [np.nan, 1.0]])
    t = np.array([[1.0, 0.0], [0.0, 1.0]])
    net.losses, net.l_out = [lo], [type("O", (), {"output": p})()]
    with pytest.raises(ValueError, match="NaNs in the loss function"):
        net.compute_losses([t])
