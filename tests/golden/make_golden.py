#!/usr/bin/env python
"""Generate the golden fixtures from the REAL reference (build container only).

Run:  python tests/golden/make_golden.py
Needs /root/reference (read-only); nothing at test/bench time does.  Writes
small .npz files next to this script:

  ref_unit_kat.npz   the reference's own unit-test vectors, captured by hooking
                     the layer classes while running tests/coverage/test_layers.py
                     and test_losses.py unmodified (they must pass)
  layer_cases.npz    seeded differential cases for every §8a row (cases.py)
  loss_cases.npz, opt_cases.npz
  net_<config>.npz   small variants of the five BASELINE configs: per-step
                     losses, first-step grads, final params (topologies.py)

It also asserts, on the spot, that the oracle reproduces every record — the
oracle is pinned here against the reference itself, and again (without the
reference) by tests/test_oracle.py against these files.
"""
import json
import os
import sys
import types
import unittest
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

# dnpy/utils.py:5 imports matplotlib (absent here): two-function stub.
_mpl = types.ModuleType("matplotlib")
_plt = types.ModuleType("matplotlib.pyplot")
_plt.imshow = lambda *a, **k: None
_plt.show = lambda *a, **k: None
_mpl.pyplot = _plt
sys.modules.setdefault("matplotlib", _mpl)
sys.modules.setdefault("matplotlib.pyplot", _plt)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore", category=DeprecationWarning)

import dnpy.layers as RL                      # noqa: E402
from dnpy import net as rnet, optimizers as ropt, losses as rlosses, metrics as rmetrics  # noqa: E402
from dnpy import initializers as rinit, regularizers as rreg  # noqa: E402

from tests.golden import cases, drive         # noqa: E402
from tests import topologies                  # noqa: E402
from oracle.oracle_net import OracleNet       # noqa: E402
from oracle import dnpy_oracle as O           # noqa: E402

REF_NS = types.SimpleNamespace(L=RL, Net=rnet.Net, opt=ropt, losses=rlosses, metrics=rmetrics,
                               init=rinit, reg=rreg)


def close(a, b, tol=1e-9):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(1.0, float(np.max(np.abs(b))) if b.size else 1.0)
    err = float(np.max(np.abs(a - b))) / scale if b.size else 0.0
    assert err <= tol, err


# ----------------------------------------------------------------------------
# 1. the reference's own unit tests, recorded
# ----------------------------------------------------------------------------
def record_reference_unit_tests():
    records = []

    def hook(cls):
        orig = cls.backward

        def backward(self):
            orig(self)
            cfg = dict(kind=type(self).__name__, in_shape=list(self.parents[0].oshape))
            for a in ("filters", "kernel_size", "strides", "padding", "pool_size", "stable"):
                if hasattr(self, a):
                    v = getattr(self, a)
                    cfg[a] = list(v) if isinstance(v, (tuple, list)) else v
            rec = dict(cfg=cfg, x=np.array(self.parents[0].output, dtype=float), out=np.array(self.output, dtype=float),
                       delta=np.array(self.delta, dtype=float), dx=np.array(self.parents[0].delta, dtype=float))
            for k, v in self.params.items():
                rec["p_" + k] = np.array(v, dtype=float)
            records.append(rec)
        cls.backward = backward
        return orig

    hooked = [(c, hook(c)) for c in (RL.Conv2D, RL.DepthwiseConv2D, RL.MaxPool2D, RL.AvgPool2D, RL.Softmax)]
    loss_recs = []
    orig_delta = rlosses.BinaryCrossEntropy.compute_delta

    def bce_delta(self, p, t):
        d = orig_delta(self, p, t)
        loss_recs.append(dict(p=np.array(p), t=np.array(t), loss=self.compute_loss(p, t), delta=np.array(d)))
        return d
    rlosses.BinaryCrossEntropy.compute_delta = bce_delta

    suite = unittest.TestSuite()
    loader = unittest.TestLoader()
    import importlib.util
    for fn in ("test_layers.py", "test_losses.py"):
        spec = importlib.util.spec_from_file_location("ref_" + fn[:-3], os.path.join(REF, "tests/coverage", fn))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        suite.addTests(loader.loadTestsFromModule(mod))
    res = unittest.TextTestRunner(verbosity=0).run(suite)
    assert res.wasSuccessful(), "reference unit tests must pass before recording them"
    for c, o in hooked:
        c.backward = o
    rlosses.BinaryCrossEntropy.compute_delta = orig_delta

    out = {"n": np.int64(len(records)), "n_loss": np.int64(len(loss_recs)),
           "tests_run": np.int64(res.testsRun)}
    for i, r in enumerate(records):
        out[f"{i}_cfg"] = np.array(json.dumps(r["cfg"]))
        for k, v in r.items():
            if k != "cfg":
                out[f"{i}_{k}"] = v
    for i, r in enumerate(loss_recs):
        for k, v in r.items():
            out[f"loss{i}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "ref_unit_kat.npz"), **out)
    print(f"ref_unit_kat: {len(records)} layer records + {len(loss_recs)} loss records from {res.testsRun} reference tests")


# ----------------------------------------------------------------------------
# 2. seeded per-operator cases
# ----------------------------------------------------------------------------
def record_layer_cases():
    out = {}
    for case in cases.LAYER_CASES:
        rec = drive.run_layer_case(REF_NS, case)
        orc = drive.run_layer_case_oracle(REF_NS, case)
        for k, v in rec.items():
            out[f"{case['name']}/{k}"] = v
            if k in orc:
                close(orc[k], v)
            elif k.startswith(("dx", "g_", "out", "idx")):
                raise AssertionError(f"oracle lacks {k} for {case['name']}")
        if case["kind"] == "MaxPool2D":        # the literal statement itself, too
            pads = drive._prepare(REF_NS, case)[3]
            l = pads
            p4 = (l.pad_top, l.pad_bottom, l.pad_left, l.pad_right)
            lit = O.maxpool2d_backward_loops(rec["delta"], rec["idx"], rec["x0"].shape, tuple(l.pool_size),
                                             tuple(l.strides), p4)
            close(lit, rec["dx0"])
    np.savez_compressed(os.path.join(HERE, "layer_cases.npz"), **out)
    print(f"layer_cases: {len(cases.LAYER_CASES)} cases, oracle matches all")

    out = {}
    for case in cases.LOSS_CASES:
        rec = drive.run_loss_case(REF_NS, case)
        for k, v in rec.items():
            out[f"{case['name']}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "loss_cases.npz"), **out)
    out = {}
    for case in cases.OPT_CASES:
        rec = drive.run_opt_case(REF_NS, case)
        for k, v in rec.items():
            out[f"{case['name']}/{k}"] = v
    np.savez_compressed(os.path.join(HERE, "opt_cases.npz"), **out)
    print("loss_cases, opt_cases written")


# ----------------------------------------------------------------------------
# 3. whole-net cases (small variants of the five configs)
# ----------------------------------------------------------------------------
def run_ref_net(name, kw, batch, steps):
    np.random.seed(42)
    net = topologies.BUILDERS[name](REF_NS, **kw)
    rs = np.random.RandomState(1234)
    x, y = topologies.make_data(name, batch * steps, rs, **kw)
    init_params = net.get_params()
    onet = OracleNet.mirror(net)
    net.set_mode("train")
    onet.set_mode("train")
    out = {"x": x, "y": y}
    for li, p in enumerate(init_params):
        for k, v in p.items():
            out[f"init/{li}/{k}"] = np.array(v)
    losses = []
    np.random.seed(777)
    for s in range(steps):
        xb, yb = [x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]]
        net.feed_input(xb)
        net.forward()
        lo, de = net.compute_losses(y_target=yb)
        net.reset_grads()
        net.do_delta(de)
        net.backward()
        if s == 0:
            for li, l in enumerate(net.fts_layers):
                for k, g in l.grads.items():
                    out[f"grad0/{li}/{k}"] = np.array(g)
                if l.delta is not None:
                    out[f"delta0/{li}"] = np.array(l.delta)
                out[f"out0/{li}"] = np.array(l.output)
        net.apply_grads()
        losses.append(lo[0])
    out["losses"] = np.array(losses)
    for li, l in enumerate(net.fts_layers):
        for k, v in l.params.items():
            out[f"final/{li}/{k}"] = np.array(v)
    # oracle, same seeds
    np.random.seed(777)
    olosses = []
    for s in range(steps):
        xb, yb = [x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]]
        olosses.append(onet.step(xb, yb)[0])
        if s == 0:
            for li, n in enumerate(onet.fts):
                for k, g in n.grads.items():
                    close(g, out[f"grad0/{li}/{k}"])
    close(np.array(olosses), out["losses"])
    for li, n in enumerate(onet.fts):
        for k, v in n.params.items():
            close(v.reshape(out[f"final/{li}/{k}"].shape), out[f"final/{li}/{k}"])
    return out


def record_nets():
    for name, (kw, batch, steps) in topologies.GOLDEN_NETS.items():
        out = run_ref_net(name, kw, batch, steps)
        np.savez_compressed(os.path.join(HERE, f"net_{name}.npz"), **out)
        print(f"net_{name}: {steps} steps, losses {np.round(out['losses'], 5)} — oracle matches")


if __name__ == "__main__":
    record_reference_unit_tests()
    record_layer_cases()
    record_nets()
