"""CPU tests: the oracle against every fixture recorded from the real reference.

* ref_unit_kat.npz — the reference's OWN unit-test vectors (tests/coverage/test_layers.py,
  test_losses.py), captured verbatim by tests/golden/make_golden.py;
* layer_cases.npz / loss_cases.npz / opt_cases.npz — seeded differential cases per §8a row;
* net_*.npz — small variants of the five BASELINE configs (losses, first-step grads, final params).

The host-side logic of dnpy_b200 (constructors, shapes, paddings, initialiser RNG order, topological
sort) is exercised on the way: the oracle mirrors nets BUILT with the product's classes.
"""
import json

import numpy as np
import pytest

from oracle import dnpy_oracle as O
from oracle.oracle_net import OracleNet
from tests import topologies
from tests.conftest import rel_err
from tests.golden import cases, drive

TOL = 1e-9


def test_reference_unit_kats(golden):
    g = golden("ref_unit_kat.npz")
    assert int(g["tests_run"]) == 13 and int(g["n"]) == 11
    for i in range(int(g["n"])):
        cfg = json.loads(str(g[f"{i}_cfg"]))
        x, out, delta, dx = g[f"{i}_x"], g[f"{i}_out"], g[f"{i}_delta"], g[f"{i}_dx"]
        kind = cfg["kind"]
        if kind in ("Conv2D", "DepthwiseConv2D"):
            w, b = g[f"{i}_p_w1"], g[f"{i}_p_b1"]
            k = cfg["kernel_size"][-2:]
            oshape = O.get_output(cfg["in_shape"][1:], k, cfg["strides"], cfg["padding"])
            pads = O.get_padding(cfg["padding"], cfg["in_shape"][1:], oshape, k, cfg["strides"])
            (pt, pb), (pl, pr) = O.get_side_paddings(pads)
            osh = (w.shape[0], *oshape)
            if kind == "Conv2D":
                o = O.conv2d_forward(x, w, b, cfg["strides"], (pt, pb, pl, pr), osh)
                d, _, _ = O.conv2d_backward(x, w, b, delta, cfg["strides"], (pt, pb, pl, pr))
                o2 = O.conv2d_forward_loops(x, w, b, cfg["strides"], (pt, pb, pl, pr), osh)
                assert rel_err(o2, out) <= TOL
            else:
                o = O.dwconv2d_forward(x, w, b, cfg["strides"], (pt, pb, pl, pr), osh)
                d, _, _ = O.dwconv2d_backward(x, w, b, delta, cfg["strides"], (pt, pb, pl, pr))
        elif kind in ("MaxPool2D", "AvgPool2D"):
            k = cfg["pool_size"]
            oshape = O.get_output(cfg["in_shape"][1:], k, cfg["strides"], cfg["padding"])
            pads = O.get_padding(cfg["padding"], cfg["in_shape"][1:], oshape[1:], k, cfg["strides"])
            (pt, pb), (pl, pr) = O.get_side_paddings(pads)
            osh = (x.shape[1], *oshape)
            if kind == "MaxPool2D":
                o, idx = O.maxpool2d_forward(x, k, cfg["strides"], (pt, pb, pl, pr), osh)
                d = O.maxpool2d_backward(delta, idx, x.shape, k, cfg["strides"], (pt, pb, pl, pr))
            else:
                o = O.avgpool2d_forward(x, k, cfg["strides"], (pt, pb, pl, pr), osh)
                d = O.avgpool2d_backward(delta, x.shape, k, cfg["strides"], (pt, pb, pl, pr))
        elif kind == "Softmax":
            o = O.softmax_forward(x, cfg["stable"])
            d = O.softmax_backward(o, delta, False)
        else:
            raise AssertionError(kind)
        assert rel_err(o, out) <= TOL, (i, kind)
        assert rel_err(d, dx) <= TOL, (i, kind)
    # losses KAT (test_losses.py:14-47)
    p, t = g["loss0_p"], g["loss0_t"]
    assert abs(O.bce_loss(p, t) - float(g["loss0_loss"])) <= 1e-12
    assert round(O.bce_loss(p, t), 4) == 0.4564
    assert rel_err(O.bce_delta(p, t), g["loss0_delta"]) <= TOL


def test_get_output_kats():
    """tests/coverage/test_layers.py:13-30."""
    cases_ = [((4, 4), (3, 3), (1, 1), (0, 0), (2, 2)), ((5, 5), (4, 4), (1, 1), (2, 2), (6, 6)),
              ((5, 5), (3, 3), (1, 1), (1, 1), (5, 5)), ((5, 5), (3, 3), (1, 1), (2, 2), (7, 7)),
              ((5, 5), (3, 3), (2, 2), (0, 0), (2, 2)), ((5, 5), (3, 3), (2, 2), (1, 1), (3, 3))]
    from dnpy_b200 import utils
    for i, k, s, p, want in cases_:
        assert tuple(O.get_output(i, k, s, p)) == want
        assert tuple(utils.get_output(input_size=i, kernel_size=k, strides=s, padding=p)) == want


@pytest.mark.parametrize("case", cases.LAYER_CASES, ids=[c["name"] for c in cases.LAYER_CASES])
def test_layer_cases(case, golden, ns):
    g = golden("layer_cases.npz")
    rec = drive.run_layer_case_oracle(ns, case)
    keys = [k.split("/", 1)[1] for k in g.files if k.startswith(case["name"] + "/")]
    assert keys
    checked = 0
    for k in keys:
        want = g[f"{case['name']}/{k}"]
        if k.startswith(("x", "pin_", "delta")):
            continue
        if k.startswith("pout_"):
            assert rel_err(np.asarray(rec[k]).reshape(want.shape), want) <= TOL, k
        elif k == "idx":
            assert np.array_equal(rec[k], want)
        else:
            assert rel_err(rec[k], want) <= TOL, k
        checked += 1
    assert checked >= 1


@pytest.mark.parametrize("case", cases.LOSS_CASES, ids=[c["name"] for c in cases.LOSS_CASES])
def test_loss_cases(case, golden):
    g = golden("loss_cases.npz")
    p, t = g[f"{case['name']}/p"], g[f"{case['name']}/t"]
    if case["kind"] == "CrossEntropy":
        loss, delta = O.ce_loss(p, t), O.ce_delta(p, t, case.get("softmax_output", False))
    elif case["kind"] == "BinaryCrossEntropy":
        loss, delta = O.bce_loss(p, t), O.bce_delta(p, t)
    elif case["kind"] in ("RMSE", "MAE", "NLL", "Hinge"):
        k = case["kind"].lower()
        loss, delta = getattr(O, k + "_loss")(p, t), getattr(O, k + "_delta")(p, t)
    else:
        loss, delta = O.mse_loss(p, t), O.mse_delta(p, t)
    assert abs(loss - float(g[f"{case['name']}/loss"])) <= 1e-12
    assert rel_err(delta, g[f"{case['name']}/delta"]) <= TOL


@pytest.mark.parametrize("case", cases.OPT_CASES, ids=[c["name"] for c in cases.OPT_CASES])
def test_opt_cases(case, golden):
    g = golden("opt_cases.npz")
    n = case["name"]
    params = {k: g[f"{n}/p0_{k}"].copy() for k in ("w1", "b1", "frozen_stat")}
    state = {k: dict(V=np.zeros_like(v), S=np.zeros_like(v)) for k, v in params.items()}
    for step in range(3):
        for k in ("w1", "b1"):
            gr = g[f"{n}/g{step}_{k}"]
            if case["kind"] == "Adam":
                O.adam_apply(params[k], gr, state[k]["V"], state[k]["S"], **case["kw"])
            elif case["kind"] == "RMSProp":
                O.rmsprop_apply(params[k], gr, state[k]["S"], **case["kw"])
            else:
                O.sgd_apply(params[k], gr, state[k]["V"], **case["kw"])
    for k in params:
        assert rel_err(params[k], g[f"{n}/p_{k}"]) <= TOL


@pytest.mark.parametrize("name", list(topologies.GOLDEN_NETS))
def test_net_cases(name, golden, ns):
    """Build with the PRODUCT's host code (same seed), mirror into the oracle, replay the recorded
    steps: initial weights must be bit-identical, losses/grads/final params equal to the reference."""
    kw, batch, steps = topologies.GOLDEN_NETS[name]
    g = golden(f"net_{name}.npz")
    np.random.seed(42)
    net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net)
    for li, n in enumerate(onet.fts):
        for k, v in n.params.items():
            assert np.array_equal(v, g[f"init/{li}/{k}"]), f"initial {li}/{k} differs from the reference"
    onet.set_mode("train")
    x, y = g["x"], g["y"]
    np.random.seed(777)
    losses = []
    for s in range(steps):
        xb, yb = [x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]]
        losses.append(onet.step(xb, yb)[0])
        if s == 0:
            for li, n in enumerate(onet.fts):
                for k, gr in n.grads.items():
                    assert rel_err(gr, g[f"grad0/{li}/{k}"]) <= TOL, (li, k)
    assert rel_err(np.array(losses), g["losses"]) <= TOL
    for li, n in enumerate(onet.fts):
        for k, v in n.params.items():
            want = g[f"final/{li}/{k}"]
            # floor 1e-3: the Dense bias in front of a BatchNorm has a mathematically-zero gradient; its
            # value is pure round-off (1e-14) that depends on the BLAS summation order of the run
            assert rel_err(v.reshape(want.shape), want, floor=1e-3) <= TOL, (li, k)


def test_loop_variants_match_vectorised():
    """The reference-loop-structured functions used for the CPU baseline equal the vectorised ones."""
    rs = np.random.RandomState(3)
    x, w, b = rs.randn(3, 2, 7, 6), rs.randn(4, 2, 3, 3), rs.randn(4)
    pads, st = (1, 1, 1, 1), (1, 1)
    osh = (4, 7, 6)
    o1, o2 = O.conv2d_forward(x, w, b, st, pads, osh), O.conv2d_forward_loops(x, w, b, st, pads, osh)
    assert rel_err(o1, o2) <= 1e-12
    d = rs.randn(*o1.shape)
    r1 = O.conv2d_backward(x, w, b, d, st, pads, (0.01, 0.02), (0.0, 0.03))
    r2 = O.conv2d_backward_loops(x, w, b, d, st, pads, (0.01, 0.02), (0.0, 0.03))
    for a, c in zip(r1, r2):
        assert rel_err(a, c) <= 1e-12
    xm = rs.randn(5, 3, 7, 7)
    out, idx = O.maxpool2d_forward(xm, (3, 3), (2, 2), (0, 0, 0, 0), (3, 3, 3))
    dm = rs.randn(*out.shape)
    a = O.maxpool2d_backward(dm, idx, xm.shape, (3, 3), (2, 2), (0, 0, 0, 0), "literal")
    c = O.maxpool2d_backward_loops(dm, idx, xm.shape, (3, 3), (2, 2), (0, 0, 0, 0))
    assert rel_err(a, c) <= 1e-12
    # per-sample == literal applied one sample at a time (the only regime the reference's goldens cover)
    ps = O.maxpool2d_backward(dm, idx, xm.shape, (3, 3), (2, 2), (0, 0, 0, 0), "per_sample")
    one = np.concatenate([O.maxpool2d_backward_loops(dm[i:i + 1], idx[i:i + 1], xm[i:i + 1].shape, (3, 3), (2, 2),
                                                     (0, 0, 0, 0)) for i in range(5)])
    assert rel_err(ps, one) <= 1e-12
