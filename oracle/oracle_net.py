"""Whole-step oracle: mirrors ``Net.fit``'s inner loop (net.py:174-195) on top of
``dnpy_oracle``'s functions — TEST INFRASTRUCTURE ONLY.

``OracleNet.mirror(net)`` walks a *built* net (either the real reference's
``dnpy.net.Net`` — used by tests/golden/make_golden.py to pin this file — or the
product's ``dnpy_b200.net.Net`` — used by the GPU parity tests) and copies its
topology, hyper-parameters and current parameter values into NumPy float64
state.  It only reads attributes (duck typing); it never calls the net's
compute, so the product path cannot route through it.
"""
import numpy as np

from . import dnpy_oracle as O


_MERGE_OPS = {"Add": np.add, "Subtract": np.subtract, "Multiply": np.multiply, "Divide": np.divide, "Power": np.power,
              "Maximum": np.maximum, "Minimum": np.minimum}


def _reg(r):
    """(l1, l2) of a dnpy regularizer object (regularizers.py:16-53)."""
    if r is None:
        return (0.0, 0.0)
    if hasattr(r, "l1_reg"):
        return (float(r.l1_reg.lmda_l1), float(r.l2_reg.lmda_l2))
    return (float(getattr(r, "lmda_l1", 0.0)), float(getattr(r, "lmda_l2", 0.0)))


def _np(v):
    return np.array(np.asarray(v), dtype=np.float64)


class _Node:
    def __init__(self, src):
        self.kind = type(src).__name__
        self.src = src
        self.parents = []
        self.output = None
        self.delta = None
        self.params = {k: _np(v) for k, v in src.params.items()}
        self.grads = {k: np.zeros_like(self.params[k]) for k in src.grads.keys()}
        self.state = {}          # optimizer state, one dict per slot
        self.cache = {}
        self.training = False
        self.frozen = bool(getattr(src, "frozen", False))


class OracleNet:
    def __init__(self):
        self.fts = []
        self.bts = []
        self.l_in = []
        self.l_out = []
        self.losses = []
        self.opt = None
        self.maxpool_route = "literal"
        self.rng = np.random

    # -- construction ------------------------------------------------------
    @classmethod
    def mirror(cls, net, maxpool_route="literal"):
        self = cls()
        self.maxpool_route = maxpool_route
        nodes = {}
        for l in net.fts_layers:
            nodes[id(l)] = _Node(l)
        for l in net.fts_layers:
            nodes[id(l)].parents = [nodes[id(p)] for p in l.parents]
        self.fts = [nodes[id(l)] for l in net.fts_layers]
        self.bts = [nodes[id(l)] for l in net.bts_layers]
        self.l_in = [nodes[id(l)] for l in net.l_in]
        self.l_out = [nodes[id(l)] for l in net.l_out]
        self.losses = [(type(lo).__name__, bool(getattr(lo, "softmax_output", False))) for lo in net.losses]
        o = net.optimizer
        self.opt = dict(kind=type(o).__name__, lr=float(o.lr))
        if self.opt["kind"] == "SGD":
            self.opt.update(momentum=float(o.momentum), nesterov=bool(o.nesterov))
        elif self.opt["kind"] == "RMSProp":
            self.opt.update(rho=float(o.rho), eps=float(o.epsilon))
        elif self.opt["kind"] == "Adam":
            self.opt.update(beta_1=float(o.beta_1), beta_2=float(o.beta_2), eps=float(o.epsilon))
        else:
            raise NotImplementedError(self.opt["kind"])
        for n in self.fts:
            # optimizers.py:27-29,69-71,101-104: state for EVERY params key (Q14)
            n.state = {k: dict(V=np.zeros_like(v), S=np.zeros_like(v)) for k, v in n.params.items()}
        return self

    def set_mode(self, mode):
        for n in self.fts:
            n.training = (mode == "train")

    # -- forward -----------------------------------------------------------
    def _forward_node(self, n):
        s, k = n.src, n.kind
        x = n.parents[0].output if n.parents else None
        P = n.params
        if k == "Input":
            return
        if k == "Dense":
            n.output = O.dense_forward(x, P['w1'], P['b1'])
        elif k in ("Reshape", "Flatten"):
            shape = (-1, *s.oshape) if not s.include_batch else s.oshape_batch
            n.output = np.reshape(x, shape)
        elif k in ("Conv2D", "PointwiseConv2D"):
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            n.output = O.conv2d_forward(x, P['w1'], P['b1'], tuple(s.strides), pads, tuple(s.oshape))
        elif k == "DepthwiseConv2D":
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            n.output = O.dwconv2d_forward(x, P['w1'], P['b1'], tuple(s.strides), pads, tuple(s.oshape))
        elif k in ("MaxPool2D", "GlobalMaxPool2D"):
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            n.output, n.cache['idx'] = O.maxpool2d_forward(x, tuple(s.pool_size), tuple(s.strides), pads, tuple(s.oshape))
        elif k in ("AvgPool2D", "GlobalAvgPool2D"):
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            n.output = O.avgpool2d_forward(x, tuple(s.pool_size), tuple(s.strides), pads, tuple(s.oshape))
        elif k in ("LeakyRelu", "Relu"):
            n.output = O.leakyrelu_forward(x, s.alpha)
        elif k == "PRelu":
            n.output = O.prelu_forward(x, P['alpha'])
        elif k == "Sigmoid":
            n.output = O.sigmoid_forward(x)
        elif k == "Tanh":
            n.output = O.tanh_forward(x)
        elif k == "Softmax":
            n.output = O.softmax_forward(x, s.stable)
        elif k == "LogSoftmax":
            n.output = O.logsoftmax_forward(x, s.stable)
        elif k == "Dropout":
            if n.training:
                n.cache['gate'] = O.dropout_mask(x.shape, s.rate, self.rng)
            n.output = O.dropout_forward(x, s.rate, n.training, n.cache.get('gate'))
        elif k == "GaussianNoise":              # regularization.py:39-47
            n.output = O.gaussian_noise_forward(x, s.mean, s.stddev, n.training, self.rng)
        elif k == "BatchNorm":
            if getattr(s, "bias_correction", False):
                raise NotImplementedError("BatchNorm(bias_correction=True) is out of scope")
            n.output, n.cache['bn'], P['moving_mu'], P['moving_var'] = O.batchnorm_forward(
                x, P['gamma'], P['beta'], P['moving_mu'], P['moving_var'], s.momentum, n.training)
        elif k in _MERGE_OPS:                   # merge.py:13-16: left fold with the layer's NumPy operator
            out = np.array(n.parents[0].output, dtype=np.float64)
            for p in n.parents[1:]:
                out = _MERGE_OPS[k](out, p.output)
            n.output = out
        elif k == "Average":                    # merge.py:71-75
            n.output = np.sum([p.output for p in n.parents], axis=0) / len(n.parents)
        elif k == "Embedding":
            n.output = O.embedding_forward(np.asarray(x).astype(np.int64), P['w1'])
        elif k == "SimpleRNN":
            n.cache['states'] = O.rnn_forward(x, P['waa'], P['wax'], P['ba'])
            n.output = n.cache['states'][:, -1, :]
        else:
            raise NotImplementedError(k)

    def forward(self):
        for n in self.fts:
            self._forward_node(n)

    # -- backward ----------------------------------------------------------
    def _backward_node(self, n):
        s, k = n.src, n.kind
        P, G = n.params, n.grads
        par = n.parents[0] if n.parents else None
        x = par.output if par is not None else None
        d = n.delta
        if k == "Input":
            return
        if k == "Dense":
            dx, gw, gb = O.dense_backward(x, P['w1'], P['b1'], d, _reg(s.kernel_regularizer), _reg(s.bias_regularizer))
            par.delta = dx
            G['w1'] += gw
            G['b1'] += gb
        elif k in ("Reshape", "Flatten"):
            shape = (-1, *s.parents[0].oshape) if not s.include_batch else s.oshape_batch
            par.delta = np.reshape(d, shape)
        elif k in ("Conv2D", "PointwiseConv2D", "DepthwiseConv2D"):
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            fn = O.dwconv2d_backward if k == "DepthwiseConv2D" else O.conv2d_backward
            dx, gw, gb = fn(x, P['w1'], P['b1'], d, tuple(s.strides), pads,
                            _reg(s.kernel_regularizer), _reg(s.bias_regularizer))
            par.delta = dx
            G['w1'] += gw
            G['b1'] += gb
        elif k in ("MaxPool2D", "GlobalMaxPool2D"):
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            par.delta = O.maxpool2d_backward(d, n.cache['idx'], x.shape, tuple(s.pool_size), tuple(s.strides),
                                             pads, self.maxpool_route)
        elif k in ("AvgPool2D", "GlobalAvgPool2D"):
            pads = (s.pad_top, s.pad_bottom, s.pad_left, s.pad_right)
            par.delta = O.avgpool2d_backward(d, x.shape, tuple(s.pool_size), tuple(s.strides), pads)
        elif k in ("LeakyRelu", "Relu"):
            par.delta = O.leakyrelu_backward(x, d, s.alpha)
        elif k == "PRelu":
            par.delta, ga = O.prelu_backward(x, d, P['alpha'])
            G['alpha'] += ga
        elif k == "Sigmoid":
            par.delta = O.sigmoid_backward(n.output, d)
        elif k == "Tanh":
            par.delta = O.tanh_backward(n.output, d)
        elif k == "Softmax":
            par.delta = O.softmax_backward(n.output, d, bool(s.ce_loss))
        elif k == "LogSoftmax":
            par.delta = d
        elif k == "Dropout":
            par.delta = O.dropout_backward(d, n.cache['gate'])
        elif k == "GaussianNoise":
            par.delta = np.array(d)
        elif k == "BatchNorm":
            par.delta, gg, gb = O.batchnorm_backward(d, P['gamma'], n.cache['bn'])
            G['gamma'] += gg.reshape(G['gamma'].shape)
            G['beta'] += gb.reshape(G['beta'].shape)
        elif k in _MERGE_OPS:
            for p in n.parents:          # merge.py:18-20: same delta to every parent (Q10)
                p.delta = d
        elif k == "Average":             # merge.py:77-79
            for p in n.parents:
                p.delta = d / len(n.parents)
        elif k == "Embedding":
            O.embedding_backward_inplace(G['w1'], np.asarray(x).astype(np.int64), d)
        elif k == "SimpleRNN":
            dx, gwaa, gwax, gba = O.rnn_backward(x, n.cache['states'], d, P['waa'], P['wax'], s.bptt_truncate)
            par.delta = dx
            G['waa'] += gwaa
            G['wax'] += gwax
            G['ba'] += gba
        else:
            raise NotImplementedError(k)

    def backward(self):
        for n in self.bts:
            self._backward_node(n)

    # -- loss / optimizer --------------------------------------------------
    def compute_losses(self, y_target):
        losses, deltas = [], []
        for i, (kind, sm) in enumerate(self.losses):
            p, t = self.l_out[i].output, np.asarray(y_target[i], dtype=np.float64)
            if kind == "CrossEntropy":
                losses.append(O.ce_loss(p, t)); deltas.append(O.ce_delta(p, t, sm))
            elif kind == "BinaryCrossEntropy":
                losses.append(O.bce_loss(p, t)); deltas.append(O.bce_delta(p, t))
            elif kind == "MSE":
                losses.append(O.mse_loss(p, t)); deltas.append(O.mse_delta(p, t))
            elif kind in ("RMSE", "MAE", "NLL", "Hinge"):
                losses.append(getattr(O, kind.lower() + "_loss")(p, t)); deltas.append(getattr(O, kind.lower() + "_delta")(p, t))
            else:
                raise NotImplementedError(kind)
        return losses, deltas

    def apply_grads(self):
        o = self.opt
        for n in self.bts:
            if n.frozen:
                continue
            for k in n.params.keys():
                if k not in n.grads:
                    continue
                st = n.state[k]
                if o["kind"] == "Adam":
                    O.adam_apply(n.params[k], n.grads[k], st['V'], st['S'], o["lr"], o["beta_1"], o["beta_2"], o["eps"])
                elif o["kind"] == "RMSProp":
                    O.rmsprop_apply(n.params[k], n.grads[k], st['S'], o["lr"], o["rho"], o["eps"])
                else:
                    O.sgd_apply(n.params[k], n.grads[k], st['V'], o["lr"], o["momentum"], o["nesterov"])

    def forward_loss(self, x_mb, y_mb):
        """feed_input + forward + losses only (what Net.evaluate does per batch, net.py:237-253)."""
        for j, n in enumerate(self.l_in):
            n.output = np.asarray(x_mb[j]) if self.l_in[j].src.__class__.__name__ == "Input" else x_mb[j]
        self.forward()
        return self.compute_losses(y_mb)[0]

    def step(self, x_mb, y_mb, apply=True):
        """One iteration of net.py:176-195; returns the list of losses."""
        for j, n in enumerate(self.l_in):
            n.output = np.asarray(x_mb[j]) if self.l_in[j].src.__class__.__name__ == "Input" else x_mb[j]
        self.forward()
        losses, deltas = self.compute_losses(y_mb)
        for n in self.fts:
            for g in n.grads.values():
                g.fill(0.0)
        for i, n in enumerate(self.l_out):
            n.delta = deltas[i]
        self.backward()
        if apply:
            self.apply_grads()
        return losses
