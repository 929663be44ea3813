"""GPU parity of whole training steps (-m gpu): the five BASELINE configs through Net.build /
feed_input / forward / compute_losses / reset_grads / backward / apply_grads and through Net.fit.

* small variants against the fixtures recorded from the real reference (tests/golden/net_*.npz);
* config-sized variants against the oracle, including loss-curve agreement over 100 steps.
Tolerances: losses rel 1e-5; first-step gradients rel 1e-5; parameters after several Adam steps
rel 1e-4 (Adam's 1/sqrt(S+eps) amplifies fp32 round-off of tiny gradients).
"""
import numpy as np
import pytest

from oracle.oracle_net import OracleNet
from tests import topologies
from tests.conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_mode():
    import dnpy_b200
    dnpy_b200.set_math("fp32")
    dnpy_b200.set_maxpool_route("literal")
    yield


def _manual_step(net, xb, yb):
    net.feed_input(xb)
    net.forward()
    lo, de = net.compute_losses(y_target=yb)
    net.reset_grads()
    net.do_delta(de)
    net.backward()
    return lo


def _forced_step(net, onet, xb, yb, max_flip):
    """One step of product and oracle from the SAME parameters with the oracle's pooling argmax teacher-forced into the
    product before its backward pass: forward + loss on the GPU, the whole step in the oracle, then every MaxPool2D's
    ``output_idxs`` is replaced by the oracle's (after asserting that at most ``max_flip`` of the entries differ — near-ties
    below the path's resolution), then the GPU backward.  With the routing identical, EVERY gradient of EVERY layer is
    comparable at every step; no layer is skipped.  Returns (gpu loss, oracle loss, largest flipped fraction)."""
    net.feed_input(xb)
    net.forward()
    return _forced_step_after_forward(net, onet, xb, yb, max_flip)


def _forced_step_after_forward(net, onet, xb, yb, max_flip):
    from dnpy_b200.darray import DArray
    lo, de = net.compute_losses(y_target=yb)
    want = onet.step(xb, yb)[0]
    worst = 0.0
    for l, n in zip(net.fts_layers, onet.fts):
        if "idx" in n.cache:
            diff = float(np.mean(np.asarray(l.output_idxs) != n.cache["idx"]))
            assert diff <= max_flip, (l.name, diff)
            worst = max(worst, diff)
            if diff > 0:
                l.output_idxs = DArray(host=n.cache["idx"], is_int=True)
    net.reset_grads()
    net.do_delta(de)
    net.backward()
    return lo[0], want, worst


@pytest.mark.parametrize("name", list(topologies.GOLDEN_NETS))
def test_small_nets_vs_reference_fixtures(name, golden, ns):
    kw, batch, steps = topologies.GOLDEN_NETS[name]
    g = golden(f"net_{name}.npz")
    np.random.seed(42)
    net = topologies.BUILDERS[name](ns, **kw)
    net.set_mode("train")
    x, y = g["x"], g["y"]
    np.random.seed(777)
    losses = []
    for s in range(steps):
        xb, yb = [x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]]
        losses.append(_manual_step(net, xb, yb)[0])
        if s == 0:
            for li, l in enumerate(net.fts_layers):
                for k, gr in l.grads.items():
                    want = g[f"grad0/{li}/{k}"]
                    # floor: a bias in front of a BatchNorm has a mathematically zero gradient (pure round-off)
                    e = rel_err(np.asarray(gr).reshape(want.shape), want, floor=1e-3)
                    assert e <= 1e-5, (li, l.name, k, e)
                if f"out0/{li}" in g.files and l.output is not None and not isinstance(l.output, np.ndarray):
                    assert rel_err(np.asarray(l.output), g[f"out0/{li}"]) <= 1e-5, (li, l.name)
        net.apply_grads()
    assert rel_err(np.array(losses), g["losses"]) <= 1e-5
    for li, l in enumerate(net.fts_layers):
        for k, v in l.params.items():
            want = g[f"final/{li}/{k}"]
            if f"grad0/{li}/{k}" in g.files and float(np.max(np.abs(g[f"grad0/{li}/{k}"]))) < 1e-12:
                continue      # gradient is pure round-off in the reference (bias in front of a BatchNorm):
                              # Adam turns 1e-18 vs 1e-9 noise into different, output-irrelevant drifts
            e = rel_err(np.asarray(v).reshape(want.shape), want, floor=1e-3)
            assert e <= 1e-4, (li, l.name, k, e)


FREE_RUNNING = [
    ("mnist_cnn", dict(f1=2, f2=4), 32, 100),
    ("mnist_mlp", dict(h1=64, h2=32), 64, 100),
]


@pytest.mark.parametrize("name,kw,batch,steps", FREE_RUNNING, ids=[c[0] for c in FREE_RUNNING])
def test_loss_curve_free_running_100_steps(name, kw, batch, steps, ns):
    """Loss-curve agreement over 100 independent steps via the public fused step (Net.train_step):
    product (fp32, GPU) and oracle (float64, CPU) start from the same weights and never exchange
    state again."""
    np.random.seed(42)
    net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net)
    net.set_mode("train")
    onet.set_mode("train")
    x, y = topologies.make_data(name, batch * 4, np.random.RandomState(1234), **kw)
    got, want = [], []
    for s in range(steps):
        lo = (s % 4) * batch
        outs = net.train_step([x[lo:lo + batch]], [y[lo:lo + batch]])
        got.append(float(outs[0][0].item()))
        want.append(onet.step([x[lo:lo + batch]], [y[lo:lo + batch]])[0])
    got, want = np.array(got), np.array(want)
    assert np.all(np.isfinite(got))
    assert rel_err(got[:10], want[:10]) <= 1e-4, rel_err(got[:10], want[:10])
    assert rel_err(got, want) <= 5e-3, rel_err(got, want)


TEACHER_FORCED = [
    ("iris_mlp", dict(), 24, 100),
    ("mnist_mlp", dict(h1=64, h2=32), 32, 30),
    ("mnist_cnn", dict(f1=8, f2=16), 16, 40),
    ("res_conv", dict(size=16, c1=8, c2=16), 8, 25),
    ("rnn_sentiment", dict(vocab=100, emb=8, seq=24, hidden=32, trunc=4), 16, 30),
]


@pytest.mark.parametrize("name,kw,batch,steps", TEACHER_FORCED, ids=[c[0] for c in TEACHER_FORCED])
def test_steps_along_the_reference_trajectory(name, kw, batch, steps, ns):
    """All five configs along the ORACLE's training trajectory.

    Training these nets is a chaotic (lr 0.1 Adam + the reference's BatchNorm backward) and
    discontinuous (pool argmax, relu gates, LITERAL max-pool routing where one flipped window moves a
    whole sample's delta) function of round-off, so a free-running fp32 run cannot shadow a float64
    run for long — see the free-running test above for the configs where it can.  Here every step
    starts from the oracle's current parameters: loss and every gradient must match at each of the
    `steps` parameter states the reference visits.  A step in which a pooling argmax differs between
    fp32 and float64 (a near-tie below fp32 resolution, <= 0.1% of the entries) has the oracle's argmax
    teacher-forced into the product before the backward pass, so every gradient of every layer is compared at
    every step; such steps must stay rare."""
    np.random.seed(42)
    net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net)
    net.set_mode("train")
    onet.set_mode("train")
    x, y = topologies.make_data(name, batch * 4, np.random.RandomState(1234), **kw)
    np.random.seed(99)
    flips = 0
    for s in range(steps):
        lo = (s % 4) * batch
        xb, yb = [x[lo:lo + batch]], [y[lo:lo + batch]]
        net.set_params([dict(n.params) for n in onet.fts])
        # Dropout: the product draws in its forward, the oracle in its own (later) forward — replay the same stream
        state = np.random.get_state()
        net.feed_input(xb)
        net.forward()
        after = np.random.get_state()
        np.random.set_state(state)
        loss, want, worst = _forced_step_after_forward(net, onet, xb, yb, 1e-3)
        np.random.set_state(after)
        assert abs(loss - want) <= 2e-5 * max(1.0, abs(want)), (s, loss, want)
        flips += worst > 0
        gmax = max([float(np.max(np.abs(g))) for n in onet.fts for g in n.grads.values()] + [1e-30])
        for li, (l, n) in enumerate(zip(net.fts_layers, onet.fts)):
            for k, g in n.grads.items():
                e = rel_err(np.asarray(l.grads[k]).reshape(g.shape), g, floor=1e-3 * gmax)
                assert e <= 1e-4, (s, l.name, k, e)
        if s in (0, steps - 1) and onet.l_in[0].delta is not None:
            # the delta of the Input layer is dead on the training path: Net skips it and materialises it on this read
            want_d = onet.l_in[0].delta
            assert rel_err(np.asarray(net.l_in[0].delta).reshape(want_d.shape), want_d) <= 1e-4, (s, "Input.delta")
    assert flips <= max(1, steps // 4), f"{flips} of {steps} steps had an fp32 argmax near-tie"


TF32_NETS = [
    ("mnist_mlp", dict(h1=512, h2=256), 128, 6),
    ("mnist_cnn", dict(f1=32, f2=64), 16, 6),
    ("res_conv", dict(size=8, c1=32, c2=64), 32, 4),
]


@pytest.mark.parametrize("math", ["tf32", "bf16"])
@pytest.mark.parametrize("name,kw,batch,steps", TF32_NETS, ids=[c[0] for c in TF32_NETS])
def test_tensor_core_steps_along_the_reference_trajectory(name, kw, batch, steps, math, ns):
    """Same protocol with math="tf32" (tcgen05 Dense/Conv kernels): the loss within rel 2e-2 of the
    float64 oracle at every visited parameter state, every gradient within 2e-2 of the largest gradient
    of its layer chain (errors of ~1e-3 per TF32 contraction accumulate through the backward chain).
    TF32 resolution (1e-3) makes pooling near-ties common: at most 1% (bf16: 3%) of the argmax entries may differ;
    the oracle's argmax is then teacher-forced into the product's pools before the backward pass (`_forced_step`), so
    the conv gradients BELOW the pools are compared at every step too (under the LITERAL routing a single flipped
    window would otherwise reroute a whole sample's delta)."""
    import dnpy_b200
    dnpy_b200.set_math(math)
    try:
        np.random.seed(42)
        net = topologies.BUILDERS[name](ns, **kw)
        onet = OracleNet.mirror(net)
        net.set_mode("train")
        onet.set_mode("train")
        x, y = topologies.make_data(name, batch * 2, np.random.RandomState(1234), **kw)
        compared = 0
        for s in range(steps):
            lo = (s % 2) * batch
            xb, yb = [x[lo:lo + batch]], [y[lo:lo + batch]]
            net.set_params([dict(n.params) for n in onet.fts])
            loss, want, _ = _forced_step(net, onet, xb, yb, 0.03 if math == "bf16" else 0.01)   # bf16 resolves 4e-3: more near-ties
            assert abs(loss - want) <= 2e-2 * max(1.0, abs(want)), (s, loss, want)
            gmax = max([float(np.max(np.abs(g))) for n in onet.fts for g in n.grads.values()] + [1e-30])
            kinds = [type(l).__name__ in ("Dense", "Conv2D", "PointwiseConv2D") for l in net.fts_layers]
            for li, (l, n) in enumerate(zip(net.fts_layers, onet.fts)):
                # 2e-2 per tensor-core contraction (BASELINE north_star): a gradient whose backward chain runs through
                # `depth` reduced-precision contractions (this layer's own and those between it and the loss) is
                # bounded by depth * 2e-2 (one more when reduced-precision layers also sit BELOW it in the forward pass)
                depth = max(1, sum(kinds[li:])) + (1 if any(kinds[:li]) else 0)    # + the forward chain below the layer
                for k, g in n.grads.items():
                    # the reference's own gradient-check metric (tests/coverage/test_gradient_checking.py:
                    # 32-92): ||a-b|| / (||a||+||b||) — norm-wise, because TF32 round-off flips a few relu
                    # gates / near-zero pre-activations and each flip moves single entries by ~1/batch
                    a = np.asarray(l.grads[k]).reshape(g.shape)
                    if float(np.max(np.abs(g))) < 1e-6 * gmax:
                        continue        # pure round-off in the float64 reference (a bias in front of a BatchNorm)
                    e = float(np.linalg.norm(a - g) / (np.linalg.norm(a) + np.linalg.norm(g)))
                    assert e <= 2e-2 * depth, (s, l.name, k, e, depth)
                    compared += 1
        n_grads = sum(len(n.grads) for n in onet.fts)
        assert compared >= steps * n_grads // 2, (compared, steps * n_grads)   # every gradient that is not pure round-off, every step
    finally:
        dnpy_b200.set_math("fp32")


def test_fit_evaluate_api(ns, capsys):
    """Net.fit / Net.evaluate end to end with the reference's print strings."""
    np.random.seed(42)
    net = topologies.mnist_cnn(ns, f1=2, f2=4)
    rs = np.random.RandomState(0)
    x, y = topologies.make_data("mnist_cnn", 32, rs)
    net.fit([x], [y], batch_size=8, epochs=2, evaluate_epoch=False, print_rate=1)
    out = capsys.readouterr().out
    assert "Epoch 1/2..." in out and "Batch 4/4 - Losses[CrossEntropy=" in out and "CategoricalAccuracy=" in out
    net.set_mode("test")
    lo, me = net.evaluate([x], [y], batch_size=8)
    assert np.isfinite(lo[0]) and 0.0 <= float(me[0][0]) <= 1.0
    # fit(print_rate=2) exercises the sync-free fused step on the non-printing epochs
    net.fit([x], [y], batch_size=8, epochs=3, evaluate_epoch=True, print_rate=2, x_test=[x], y_test=[y])


def test_evaluate_in_chunks_equals_the_literal_per_batch_loop(ns, monkeypatch):
    """Net.evaluate(batch_size=1) as Net.fit calls it (net.py:197-210): the chunked test-mode evaluation must
    return what the literal one-forward-per-batch loop returns, and what the float64 oracle computes per sample."""
    np.random.seed(7)
    net = topologies.mnist_cnn(ns, f1=2, f2=4)
    onet = OracleNet.mirror(net)
    rs = np.random.RandomState(3)
    x, y = topologies.make_data("mnist_cnn", 37, rs)
    net.set_mode("test"); onet.set_mode("test")
    monkeypatch.setattr(type(net), "EVAL_CHUNK", 16)            # several chunks, the last one ragged
    lo_c, me_c = net.evaluate([x], [y], batch_size=1)
    monkeypatch.setenv("DNPY_B200_EVAL_LITERAL", "1")
    lo_l, me_l = net.evaluate([x], [y], batch_size=1)
    assert abs(lo_c[0] - lo_l[0]) <= 1e-6 * max(1.0, abs(lo_l[0]))
    assert abs(float(me_c[0][0]) - float(me_l[0][0])) <= 1e-9
    # batch_size that does not divide the set: the reference drops the remainder
    monkeypatch.delenv("DNPY_B200_EVAL_LITERAL")
    lo_c5, _ = net.evaluate([x], [y], batch_size=5)
    monkeypatch.setenv("DNPY_B200_EVAL_LITERAL", "1")
    lo_l5, _ = net.evaluate([x], [y], batch_size=5)
    assert abs(lo_c5[0] - lo_l5[0]) <= 1e-6 * max(1.0, abs(lo_l5[0]))
    # float64 oracle, one sample at a time
    want = np.mean([onet.forward_loss([x[i:i + 1]], [y[i:i + 1]])[0] for i in range(len(x))])
    assert abs(lo_c[0] - want) <= 1e-5 * max(1.0, abs(want))


def test_fit_with_resident_dataset_equals_host_minibatches(ns, monkeypatch, capsys):
    """Net.fit uploads the dataset once and slices minibatches on the device: same parameters, same printed
    losses as feeding host minibatches step by step (also covers the last ragged evaluate chunk and a non-printing epoch)."""
    rs = np.random.RandomState(5)
    x, y = topologies.make_data("mnist_cnn", 40, rs)
    runs = []
    for no_res in ("", "1"):
        if no_res:
            monkeypatch.setenv("DNPY_B200_NO_RESIDENT", "1")
        np.random.seed(11)
        net = topologies.mnist_cnn(ns, f1=2, f2=4)
        net.fit([x], [y], batch_size=8, epochs=3, evaluate_epoch=True, print_rate=2, x_test=[x[:16]], y_test=[y[:16]])
        runs.append(([dict((k, np.asarray(v).copy()) for k, v in l.params.items()) for l in net.fts_layers], capsys.readouterr().out))
    for pa, pb in zip(runs[0][0], runs[1][0]):
        for k in pa:
            assert np.array_equal(pa[k], pb[k]), k
    assert runs[0][1] == runs[1][1]


def test_save_load_roundtrip_device(tmp_path, ns):
    np.random.seed(1)
    net = topologies.mnist_cnn(ns, f1=2, f2=4)
    rs = np.random.RandomState(0)
    x, y = topologies.make_data("mnist_cnn", 16, rs)
    net.set_mode("train")
    for _ in range(3):
        net.train_step([x], [y])
    f = str(tmp_path / "m.pkl")
    net.save(f, save_grads=True)
    np.random.seed(2)
    net2 = topologies.mnist_cnn(ns, f1=2, f2=4)
    net2.load(f)
    net.set_mode("test"); net2.set_mode("test")
    a, _ = net.evaluate([x], [y], batch_size=16)
    b, _ = net2.evaluate([x], [y], batch_size=16)
    assert abs(a[0] - b[0]) <= 1e-6


def test_full_size_properties(ns):
    """BASELINE-size tensors (B=8192, 32x26x26): size-independent properties instead of the oracle.
    * conv linearity: conv(x1 + x2) - K*b == (conv(x1) - K*b) + (conv(x2) - K*b)
    * pool: output == gather of the input at the recorded argmax; idx in range; checksum of the
      per-sample backward == checksum of the incoming delta (every delta lands exactly once)."""
    import dnpy_b200
    L = ns.L
    B = 8192
    rs = np.random.RandomState(0)
    x1, x2 = rs.rand(B, 1, 28, 28), rs.rand(B, 1, 28, 28)
    outs = []
    for xin in (x1, x2, x1 + x2):
        l0 = L.Input(shape=(1, 28, 28)); l0.output = xin
        np.random.seed(5)
        c = L.Conv2D(l0, 32, kernel_size=(3, 3), strides=(1, 1), padding="none")
        c.initialize()
        c.forward()
        outs.append(c.output.dev().clone())
    assert float((outs[2] - outs[0] - outs[1]).abs().max()) <= 1e-5 * float(outs[2].abs().max())
    l0 = L.Input(shape=(32, 26, 26))
    from dnpy_b200.darray import DArray
    l0.output = DArray.device(outs[2])
    p = L.MaxPool2D(l0, pool_size=(3, 3), strides=(2, 2), padding="none")
    p.forward()
    idx = p.output_idxs.dev().long()
    assert int(idx.min()) >= 0 and int(idx.max()) <= 8
    import torch
    oy = torch.arange(12, device="cuda").view(1, 1, 12, 1) * 2 + idx // 3
    ox = torch.arange(12, device="cuda").view(1, 1, 1, 12) * 2 + idx % 3
    flat = outs[2].view(B, 32, -1)
    gathered = torch.gather(flat, 2, (oy * 26 + ox).view(B, 32, -1)).view(B, 32, 12, 12)
    assert torch.equal(gathered, p.output.dev())
    dnpy_b200.set_maxpool_route("per_sample")
    d = torch.rand(B, 32, 12, 12, device="cuda")
    p.delta = DArray.device(d)
    p.backward()
    tot_in, tot_out = float(d.double().sum()), float(l0.delta.dev().double().sum())
    assert abs(tot_in - tot_out) <= 1e-6 * abs(tot_in)


def test_reference_gradient_check_protocol_through_the_drop_in(ns):
    """SURVEY 8f-3 / the reference's tests/coverage/test_gradient_checking.py: the numerical gradient-check harness
    (tests/gradcheck.py — validated against the real reference in float64 at the reference's own eps=1e-6 / 1e-6 bound)
    driven through the drop-in's get_params / set_params / get_grads / params2vector API on the two nets the reference
    checks that need no download: the iris MLP (test_iris_model1: Dense20-Relu-Dense15-Relu-Dense3-Softmax, CE) and the
    Boston-style regression net (test_boston_model1: Dense15-Tanh-Dense1, MSE; synthetic 13-feature data, the dataset is
    gone from sklearn).  fp32 arithmetic: probe eps=1e-3, bounds 1e-2 (Relu kinks: the float64 reference itself shows
    4e-3 at this eps) and 1e-3 (smooth Tanh net)."""
    import random
    from dnpy_b200 import utils
    from sklearn import datasets
    from tests.gradcheck import gradient_check
    L = ns.L
    np.random.seed(42)
    random.seed(42)
    iris = datasets.load_iris()
    X = (iris.data - iris.data.mean(0)) / iris.data.std(0)
    Y = utils.to_categorical(iris.target, num_classes=3)
    idx = np.arange(len(X)); np.random.shuffle(idx)
    X, Y = X[idx], Y[idx]
    l_in = L.Input(shape=X[0].shape)
    l = L.Relu(L.Dense(l_in, 20)); l = L.Relu(L.Dense(l, 15))
    l_out = L.Softmax(L.Dense(l, 3))
    m = ns.Net()
    m.build(l_in=[l_in], l_out=[l_out], optimizer=ns.opt.Adam(lr=10e-2), losses=[ns.losses.CrossEntropy()],
            metrics=[[ns.metrics.CategoricalAccuracy()]], debug=False)
    ok, worst = gradient_check(m, utils, [X], [Y], batch_size=15, max_error=1e-2, max_samples=25, eps=1e-3)
    assert ok, worst
    rs = np.random.RandomState(0)
    Xr = rs.randn(100, 13); Yr = Xr @ rs.randn(13, 1) + 0.1 * rs.randn(100, 1)
    l_in = L.Input(shape=(13,))
    l_out = L.Dense(L.Tanh(L.Dense(l_in, 15)), 1)
    m = ns.Net()
    m.build(l_in=[l_in], l_out=[l_out], optimizer=ns.opt.Adam(lr=10e-2), losses=[ns.losses.MSE()],
            metrics=[[ns.metrics.MSE(), ns.metrics.MAE()]], debug=False)
    ok, worst = gradient_check(m, utils, [Xr], [Yr], batch_size=10, max_error=1e-3, max_samples=25, eps=1e-3, burn_in_passes=1)
    assert ok, worst


def test_load_after_a_captured_step_keeps_graph_and_buffers_coherent(tmp_path, ns):
    """Net.load / set_params after the step graph was captured: parameters WITHOUT a gradient (BatchNorm moving
    statistics) must be overwritten in place — a captured graph bakes their addresses — so graph replays after the load
    equal an eager run from the same loaded state."""
    import dnpy_b200
    x, y = topologies.make_data("res_conv", 8 * 2, np.random.RandomState(3), size=8)
    kw = dict(size=8, c1=4, c2=8)

    def fresh(seed):
        np.random.seed(seed)
        n = topologies.res_conv(ns, **kw)
        n.set_mode("train")
        return n
    a = fresh(1)
    for s in range(5):                       # eager, eager, capture, replay, replay
        a.train_step([x[:8]], [y[:8]])
    f = str(tmp_path / "bn.pkl")
    donor = fresh(2)
    for s in range(2):
        donor.train_step([x[8:]], [y[8:]])
    donor.save(f)
    a.load(f)                                # moving_mu / moving_var come back from the pickle
    b = fresh(1)                             # the same history (Adam state included), without the graph
    b.use_graph = False
    for s in range(5):
        b.train_step([x[:8]], [y[:8]])
    b.load(f)
    for s in range(3):
        a.train_step([x[:8]], [y[:8]])       # graph replays
        b.train_step([x[:8]], [y[:8]])       # eager
    for la, lb in zip(a.fts_layers, b.fts_layers):
        for k in la.params.keys():
            if k.startswith("moving"):
                # a stale (pre-load) buffer would be O(1) off; split-K atomics make graph and eager runs differ by ~1e-5
                assert rel_err(np.asarray(la.params[k]), np.asarray(lb.params[k])) <= 1e-3, (la.name, k)
    a.set_mode("test"); b.set_mode("test")
    la_, _ = a.evaluate([x], [y], batch_size=8)
    lb_, _ = b.evaluate([x], [y], batch_size=8)
    assert abs(la_[0] - lb_[0]) <= 1e-4 * max(1.0, abs(lb_[0]))


def test_gaussian_noise_net_trains_through_the_captured_graph(ns):
    """GaussianNoise (examples/6_mnist_conv.py:48) inside Net.train_step: the noise is drawn on the host into a static
    buffer before each capture / replay, so the graph path equals the eager path draw for draw."""
    L = ns.L

    def build():
        np.random.seed(4)
        l_in = L.Input(shape=(12,))
        l = L.GaussianNoise(L.Dense(l_in, 9), mean=0.0, stddev=0.3)
        l_out = L.Softmax(L.Dense(L.Relu(l), 3))
        n = ns.Net()
        n.build(l_in=[l_in], l_out=[l_out], optimizer=ns.opt.Adam(lr=0.01), losses=[ns.losses.CrossEntropy()],
                metrics=[[ns.metrics.CategoricalAccuracy()]], debug=False, smart_derivatives=True)
        n.set_mode("train")
        return n
    rs = np.random.RandomState(2)
    x = rs.randn(16, 12); y = np.eye(3)[rs.randint(0, 3, 16)]
    outs = []
    for graph in (True, False):
        n = build()
        n.use_graph = graph
        np.random.seed(8)
        for s in range(6):
            n.train_step([x], [y])
        outs.append([np.asarray(v).copy() for l in n.fts_layers for v in l.params.values()])
    for pa, pb in zip(*outs):
        assert rel_err(pa, pb) <= 1e-6


@pytest.mark.parametrize("route", ["per_sample", "literal"])
@pytest.mark.parametrize("math", ["fp32", "bf16"])
def test_fused_conv_pool_act_block(math, route, ns):
    """The fused first block of the CNN (csrc/fused_cpa.cu: Conv2D(1->32, 3x3) -> MaxPool2D(3x3/2) -> Relu in one forward
    and one backward kernel; PER_SAMPLE pool routing and the reference's LITERAL routing — every delta into sample 0,
    last sample per argmax offset wins) against the float64 oracle with the same routing: loss, every
    gradient, and every tensor the fused kernels never write (conv.output, pool.output, pool.output_idxs, pool.delta,
    conv.delta, Input.delta — LazyDArrays computed by the layer-by-layer kernels on the read).  Step 2 replaces the pool's
    argmax by the oracle's (honoured by the fused backward: dnb_idx_gate_set); step 2 forces the layer-by-layer backward
    after a fused forward (the fallback chain through the lazies)."""
    import dnpy_b200
    from dnpy_b200.darray import DArray, LazyDArray
    dnpy_b200.set_math(math)
    dnpy_b200.set_maxpool_route(route)
    tol = 1e-4 if math == "fp32" else 2e-2
    try:
        np.random.seed(42)
        net = topologies.mnist_cnn(ns, f1=32, f2=64)
        conv, pool, act = net.fts_layers[1], net.fts_layers[2], net.fts_layers[3]
        assert conv._cpa is not None and conv._cpa[0] is pool and conv._cpa[1] is act, "block not planned"
        onet = OracleNet.mirror(net, maxpool_route=route)
        net.set_mode("train"); onet.set_mode("train")
        batch = 96
        x, y = topologies.make_data("mnist_cnn", batch * 3, np.random.RandomState(1234))
        kinds = [type(l).__name__ in ("Dense", "Conv2D", "PointwiseConv2D") for l in net.fts_layers]
        for s in range(3):
            xb, yb = [x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]]
            net.set_params([dict(n.params) for n in onet.fts])
            net.feed_input(xb)
            net.forward()
            assert isinstance(conv.output, LazyDArray) and not conv.output.materialised, "fused forward did not run"
            lo, de = net.compute_losses(y_target=yb)
            want = onet.step(xb, yb)[0]
            assert abs(lo[0] - want) <= (2e-5 if math == "fp32" else 2e-2) * max(1.0, abs(want))
            o_conv, o_pool, o_act = onet.fts[1], onet.fts[2], onet.fts[3]
            # fp32: the SIMT block; bf16: its forward runs on the tensor core with BF16 operands (csrc/tc_conv_c1p.cu)
            assert rel_err(np.asarray(act.output), o_act.output) <= (1e-5 if math == "fp32" else 2e-2)
            flips = float(np.mean(np.asarray(pool.output_idxs) != o_pool.cache["idx"]))
            assert flips <= (1e-3 if math == "fp32" else 0.03), flips
            for l, n in zip(net.fts_layers, onet.fts):      # teacher-force the oracle's argmax (one flipped window moves ~1e-3
                if "idx" in n.cache:                        # of a filter's gradient): the fused backward honours the assignment
                    l.output_idxs = DArray(host=n.cache["idx"], is_int=True)
            if s == 2:
                conv._lazy_dx = False                       # forces the layer-by-layer backward after a fused forward
            net.reset_grads()
            net.do_delta(de)
            net.backward()
            if s < 2:
                assert isinstance(conv.delta, LazyDArray) and not conv.delta.materialised, "fused backward did not run"
            else:
                conv._lazy_dx = True
            for li, (l, n) in enumerate(zip(net.fts_layers, onet.fts)):
                depth = max(1, sum(kinds[li:])) + (1 if any(kinds[:li]) else 0)
                for k, g in n.grads.items():
                    a = np.asarray(l.grads[k]).reshape(g.shape)
                    if math == "fp32":
                        e = rel_err(a, g)
                    else:       # (LITERAL routing: the last sample's delta wins and is zero below the second pool — 0 == 0)
                        e = float(np.linalg.norm(a - g) / max(np.linalg.norm(a) + np.linalg.norm(g), 1e-30))
                    assert e <= tol * (depth if math != "fp32" else 1), (s, l.name, k, e)
            if s == 1 and math == "fp32":       # everything the fused kernels skipped, on demand
                assert rel_err(np.asarray(conv.output), o_conv.output) <= 1e-5
                assert rel_err(np.asarray(pool.output), o_pool.output) <= 1e-5
                assert rel_err(np.asarray(pool.delta), o_pool.delta) <= 1e-4
                assert rel_err(np.asarray(conv.delta), o_conv.delta) <= 1e-4
                assert rel_err(np.asarray(net.l_in[0].delta), onet.l_in[0].delta) <= 1e-4
        # the same through the captured step graph: free-running product vs oracle for a few steps
        np.random.seed(42)
        net = topologies.mnist_cnn(ns, f1=32, f2=64)
        onet = OracleNet.mirror(net, maxpool_route=route)
        net.set_mode("train"); onet.set_mode("train")
        got, ref = [], []
        for s in range(6):
            outs = net.train_step([x[:batch]], [y[:batch]])
            got.append(float(outs[0][0].item()))
            ref.append(onet.step([x[:batch]], [y[:batch]])[0])
        # Adam at lr 0.01 on 32/64 filters drives the loss to ~10 within two steps (chaotic): the first steps are comparable
        assert np.all(np.isfinite(got))
        assert rel_err(np.array(got[:2]), np.array(ref[:2])) <= (1e-3 if math == "fp32" else 2e-2), (got, ref)
    finally:
        dnpy_b200.set_math("fp32")
        dnpy_b200.set_maxpool_route("literal")


@pytest.mark.parametrize("pool_geom", [((3, 3), (2, 2), 12), ((2, 2), (2, 2), 12), ((3, 3), (1, 1), 8), ((4, 4), (3, 3), 13)])
@pytest.mark.parametrize("math", ["fp32", "bf16"])
def test_fused_pool_act_pair(math, pool_geom, ns):
    """MaxPool2D -> LeakyRelu where the activation is the pool's only consumer (here behind the tcgen05 Conv2D 32 -> 64 of the
    CNN): the pool's forward kernel writes ONLY the activation's output and one byte per pooled element
    (dnb_maxpool2d_act_fwd), the backward pair is one pass from the activation's delta and the bytes (dnb_pool_act_bwd,
    per-sample routing).  Ring-kernel geometries (3x3/2, 2x2/2, 3x3/1) and a generic one (4x4/3); 600 images.  Against the
    float64 oracle: activation output, argmax, and — with the oracle's argmax teacher-forced — every gradient and delta;
    the tensors the fused kernels never wrote (pool.output, pool.output_idxs, pool.delta) come back through the lazies."""
    import dnpy_b200
    from dnpy_b200.darray import DArray, LazyDArray
    L = ns.L
    dnpy_b200.set_math(math)
    dnpy_b200.set_maxpool_route("per_sample")
    tol = 1e-5 if math == "fp32" else 2e-2
    size = pool_geom[2]

    def build(optimizer=None):
        np.random.seed(7)
        l_in = L.Input(shape=(32, size, size))
        conv = L.Conv2D(l_in, filters=64, kernel_size=(3, 3), strides=(1, 1), padding="same")
        pool = L.MaxPool2D(conv, pool_size=pool_geom[0], strides=pool_geom[1], padding="none")
        act = L.LeakyRelu(pool, alpha=0.3)
        l_out = L.Softmax(L.Dense(L.Reshape(act, shape=(-1)), 10))
        net = ns.Net()
        net.build(l_in=[l_in], l_out=[l_out], optimizer=optimizer or ns.opt.Adam(lr=0.001), losses=[ns.losses.CrossEntropy()],
                  metrics=[[ns.metrics.CategoricalAccuracy()]], debug=False, smart_derivatives=True)
        return net, conv, pool, act
    try:
        net, conv, pool, act = build()
        assert pool._pa is act and act._cpa_block is pool, "pair not planned"
        onet = OracleNet.mirror(net, maxpool_route="per_sample")
        net.set_mode("train"); onet.set_mode("train")
        batch = 600
        rs = np.random.RandomState(5)
        x = rs.randn(batch, 32, size, size)
        y = np.eye(10)[rs.randint(0, 10, batch)]
        b1 = rs.randn(64) * 0.1                        # a non-zero bias (the reference adds it K times)
        net.fts_layers[1].params['b1'] = b1
        onet.fts[1].params['b1'] = b1.copy()
        net.feed_input([x])
        net.forward()
        assert isinstance(pool.output, LazyDArray) and not pool.output.materialised, "fused forward did not run"
        # in the tensor-core modes the pair runs in the EPILOGUE of the tcgen05 convolution where the kernel takes the geometry
        # (dnb_conv2d_pool_act_fwd: the convolution output never reaches HBM either)
        in_epilogue = isinstance(conv.output, LazyDArray)
        assert in_epilogue == conv._tcp_usable(batch), (math, pool_geom, in_epilogue)
        if pool_geom[:2] == ((3, 3), (2, 2)):              # the CNN's own geometry
            assert in_epilogue == (math == "bf16")
        lo, de = net.compute_losses(y_target=[y])
        want = onet.step([x], [y])[0]
        assert abs(lo[0] - want) <= max(tol, 2e-5) * max(1.0, abs(want))
        o_conv, o_pool, o_act = onet.fts[1], onet.fts[2], onet.fts[3]
        assert rel_err(np.asarray(act.output), o_act.output) <= tol
        flips = float(np.mean(np.asarray(pool.output_idxs) != o_pool.cache["idx"]))
        assert flips <= (1e-4 if math == "fp32" else 0.03), flips
        pool.output_idxs = DArray(host=o_pool.cache["idx"], is_int=True)
        net.reset_grads()
        net.do_delta(de)
        net.backward()
        assert not pool.output.materialised, "the backward pass must not need the pooled pre-activation"
        assert not (in_epilogue and conv.output.materialised), "nor the convolution output"
        for l, n in zip(net.fts_layers, onet.fts):
            for k, g in n.grads.items():
                a = np.asarray(l.grads[k]).reshape(g.shape)
                e = float(np.linalg.norm(a - g) / max(np.linalg.norm(a) + np.linalg.norm(g), 1e-30))
                assert e <= max(tol, 1e-5), (l.name, k, e)
        if math == "fp32":       # a bf16 sign flip of a near-zero pooled value swaps delta <-> alpha*delta: compared in fp32 only
            assert rel_err(np.asarray(conv.delta), o_conv.delta) <= 1e-5       # dY rebuilt from the byte tensor
        # everything the fused forward / backward skipped, on demand (the layer-by-layer kernels)
        assert rel_err(np.asarray(pool.output), o_pool.output) <= tol
        assert rel_err(np.asarray(conv.output), o_conv.output) <= tol
        if math == "fp32":
            assert rel_err(np.asarray(pool.delta), o_pool.delta) <= 1e-5
        # the captured step graph: fused pair against the two layers run separately, free running (momentum SGD: parameter
        # differences stay proportional to gradient differences — Adam turns the round-off of a near-zero gradient into lr)
        outs = []
        for fused in (True, False):
            n2, c2, p2, a2 = build(ns.opt.SGD(lr=0.01, momentum=0.9))
            if not fused:
                p2._pa = None
                a2._cpa_block = None
                c2._tcp = None
            n2.set_mode("train")
            for s in range(4):
                n2.train_step([x[:256]], [y[:256]])
            assert (p2._bufs.get('pa_idx_gate') is not None) == fused
            outs.append([np.asarray(v).copy() for l in n2.fts_layers for v in l.params.values()])
        for pa, pb in zip(*outs):
            assert rel_err(pa, pb) <= 1e-5          # identical arithmetic up to the order of atomic partial sums in the weight gradients
    finally:
        dnpy_b200.set_math("fp32")
        dnpy_b200.set_maxpool_route("literal")
