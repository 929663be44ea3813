"""Drive one differential case through a dnpy-API namespace or through the oracle.

``run_layer_case(ns, case)`` is namespace-generic: with the real reference's
namespace it produces the golden record (make_golden.py, build container only);
with ``dnpy_b200``'s namespace it exercises the CUDA path (GPU tests).
``run_layer_case_oracle`` builds the layer with ``ns`` only for its host-side
config/shape logic and then computes with the oracle.
"""
import types

import numpy as np

from .cases import case_inputs, decode_kwargs, draw_params


def _make_layers(ns, case):
    L = ns.L
    rs, xs = case_inputs(case)
    ins = []
    for x in xs:
        l0 = L.Input(shape=tuple(case["in_shape"]))
        l0.output = x
        ins.append(l0)
    cls = getattr(L, case["kind"])
    kw = decode_kwargs(ns, case["kw"])
    l1 = cls(ins if case.get("n_inputs", 1) > 1 else ins[0], **kw)
    return rs, xs, ins, l1


def _prepare(ns, case):
    rs, xs, ins, l1 = _make_layers(ns, case)
    np.random.seed(case.get("seed", 5))
    l1.initialize()
    new_params = draw_params(rs, l1.params, case.get("pscale", 0.5))
    for k, v in new_params.items():
        l1.params[k] = v
    l1.training = bool(case.get("training", True))
    return rs, xs, ins, l1, new_params


def run_layer_case(ns, case):
    rs, xs, ins, l1, new_params = _prepare(ns, case)
    np.random.seed(case.get("seed", 5))        # Dropout draws from the global stream in forward
    l1.forward()
    out = np.array(np.asarray(l1.output), dtype=np.float64)
    rec = {"out": out}
    for i, x in enumerate(xs):
        rec[f"x{i}"] = np.asarray(x)
    for k, v in new_params.items():
        rec[f"pin_{k}"] = v
    if hasattr(l1, "output_idxs") and l1.output_idxs is not None:
        rec["idx"] = np.asarray(l1.output_idxs).astype(np.int64)
    if case.get("backward", True):
        delta = rs.randn(*out.shape)
        rec["delta"] = delta
        for k in l1.grads.keys():
            l1.grads[k].fill(0.0)
        l1.delta = delta
        l1.backward()
        for i, l0 in enumerate(ins):
            if l0.delta is not None:
                rec[f"dx{i}"] = np.array(np.asarray(l0.delta), dtype=np.float64)
        for k in l1.grads.keys():
            rec[f"g_{k}"] = np.array(np.asarray(l1.grads[k]), dtype=np.float64)
    for k in l1.params.keys():
        rec[f"pout_{k}"] = np.array(np.asarray(l1.params[k]), dtype=np.float64)
    return rec


def run_layer_case_oracle(ns, case, maxpool_route="literal"):
    """Same case, computed by the oracle (host-only; ``ns`` supplies shape logic)."""
    from oracle.oracle_net import OracleNet
    rs, xs, ins, l1, new_params = _prepare(ns, case)
    fake = types.SimpleNamespace(fts_layers=ins + [l1], bts_layers=[l1] + ins[::-1], l_in=ins, l_out=[l1],
                                 losses=[], optimizer=ns.opt.Adam())
    onet = OracleNet.mirror(fake, maxpool_route=maxpool_route)
    node = onet.l_out[0]
    node.training = bool(case.get("training", True))
    for n, x in zip(onet.l_in, xs):
        n.output = np.asarray(x)
    np.random.seed(case.get("seed", 5))
    onet._forward_node(node)
    rec = {"out": np.array(node.output, dtype=np.float64)}
    if "idx" in node.cache:
        rec["idx"] = node.cache["idx"]
    if case.get("backward", True):
        delta = rs.randn(*rec["out"].shape)
        node.delta = delta
        onet._backward_node(node)
        for i, n in enumerate(onet.l_in):
            if n.delta is not None:
                rec[f"dx{i}"] = np.array(n.delta, dtype=np.float64)
        for k, g in node.grads.items():
            rec[f"g_{k}"] = g
    for k, v in node.params.items():
        rec[f"pout_{k}"] = v
    return rec


def run_loss_case(ns, case):
    rs = np.random.RandomState(sum(map(ord, case["name"])))
    n, c = case["shape"]
    if case["kind"] == "CrossEntropy":
        z = rs.randn(n, c)
        p = np.exp(z) / np.exp(z).sum(axis=1, keepdims=True)
        t = np.eye(c)[rs.randint(0, c, n)]
    elif case["kind"] == "BinaryCrossEntropy":
        p = rs.rand(n, c) * 0.98 + 0.01
        t = rs.randint(0, 2, (n, c)).astype(float)
    elif case["kind"] == "NLL":                  # log-probabilities against one-hot targets
        z = rs.randn(n, c)
        p = z - np.log(np.exp(z).sum(axis=1, keepdims=True))
        t = np.eye(c)[rs.randint(0, c, n)]
    elif case["kind"] == "Hinge":                # +-1 targets, scores on both sides of the margin
        p = rs.randn(n, c) * 1.5
        t = rs.randint(0, 2, (n, c)).astype(float) * 2.0 - 1.0
    else:
        p = rs.randn(n, c)
        t = rs.randn(n, c)
    lo = getattr(ns.losses, case["kind"])()
    if case.get("softmax_output"):
        lo.softmax_output = True
    return dict(p=p, t=t, loss=np.float64(lo.compute_loss(p, t)),
                delta=np.array(np.asarray(lo.compute_delta(p, t)), dtype=np.float64))


def run_opt_case(ns, case, steps=3):
    rs = np.random.RandomState(sum(map(ord, case["name"])))
    params = {"w1": rs.randn(4, 3), "b1": rs.randn(1, 3), "frozen_stat": rs.randn(3)}
    p0 = {k: v.copy() for k, v in params.items()}
    opt = getattr(ns.opt, case["kind"])(**case["kw"])
    opt.initialize(params)
    gs = []
    for _ in range(steps):
        grads = {"w1": rs.randn(4, 3), "b1": rs.randn(1, 3)}
        gs.append(grads)
        opt.apply(params, grads)
    rec = {f"p0_{k}": v for k, v in p0.items()}
    rec.update({f"p_{k}": np.array(np.asarray(v), dtype=np.float64) for k, v in params.items()})
    for i, g in enumerate(gs):
        rec.update({f"g{i}_{k}": v for k, v in g.items()})
    return rec
