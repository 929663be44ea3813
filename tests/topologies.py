"""The five BASELINE.json configs as namespace-generic builders.

``ns`` is any object exposing the dnpy API surface:
  ns.L (layer classes), ns.Net, ns.opt, ns.losses, ns.metrics, ns.init, ns.reg
so the same function builds the topology in the real reference (golden
generation, build container only) and in ``dnpy_b200`` (tests, bench).
Topologies follow SURVEY.md §8d.
"""
import types

import numpy as np


def product_ns():
    import dnpy_b200.layers as L
    from dnpy_b200 import net, optimizers, losses, metrics, initializers, regularizers
    return types.SimpleNamespace(L=L, Net=net.Net, opt=optimizers, losses=losses, metrics=metrics,
                                 init=initializers, reg=regularizers)


def _build(ns, l_in, l_out, optimizer, loss, metric, smart):
    net = ns.Net()
    net.build(l_in=[l_in], l_out=[l_out], optimizer=optimizer, losses=[loss], metrics=[[metric]],
              debug=False, smart_derivatives=smart)
    return net


def iris_mlp(ns, features=4, classes=3):
    """cfg1 — examples/2_iris.py:42-62."""
    L = ns.L
    l_in = L.Input(shape=(features,))
    l = L.Dense(l_in, 20, kernel_regularizer=ns.reg.L2(lmda=0.01), bias_regularizer=ns.reg.L1(lmda=0.01))
    l = L.Relu(l)
    l = L.Dense(l, 15)
    l = L.BatchNorm(l)
    l = L.Dropout(l, 0.1)
    l = L.Relu(l)
    l = L.Dense(l, classes)
    l_out = L.Softmax(l)
    return _build(ns, l_in, l_out, ns.opt.Adam(lr=10e-2), ns.losses.CrossEntropy(),
                  ns.metrics.CategoricalAccuracy(), False)


def mnist_mlp(ns, h1=512, h2=256, classes=10, smart=True):
    """cfg2 — examples/3_mnist_mlp.py:38-54 with BASELINE's 784-512-256-10 widths."""
    L = ns.L
    l_in = L.Input(shape=(28, 28))
    l = L.Reshape(l_in, shape=(28 * 28,))
    l = L.Relu(L.Dense(l, h1))
    l = L.Relu(L.Dense(l, h2))
    l_out = L.Softmax(L.Dense(l, classes))
    return _build(ns, l_in, l_out, ns.opt.Adam(lr=0.001), ns.losses.CrossEntropy(),
                  ns.metrics.CategoricalAccuracy(), smart)


def mnist_cnn(ns, f1=2, f2=4, classes=10, size=28):
    """cfg3 — examples/6_mnist_conv.py:42-73 with its commented second block
    enabled and GaussianNoise dropped."""
    L = ns.L
    l_in = L.Input(shape=(1, size, size))
    l = L.Conv2D(l_in, filters=f1, kernel_size=(3, 3), strides=(1, 1), padding="none")
    l = L.MaxPool2D(l, pool_size=(3, 3), strides=(2, 2), padding="none")
    l = L.Relu(l)
    l = L.Conv2D(l, filters=f2, kernel_size=(3, 3), strides=(1, 1), padding="same")
    l = L.MaxPool2D(l, pool_size=(3, 3), strides=(2, 2), padding="none")
    l = L.Relu(l)
    l = L.Reshape(l, shape=(-1))
    l = L.Dense(l, classes, kernel_initializer=ns.init.RandomUniform())
    l_out = L.Softmax(l)
    return _build(ns, l_in, l_out, ns.opt.Adam(lr=0.01), ns.losses.CrossEntropy(),
                  ns.metrics.CategoricalAccuracy(), True)


def res_conv(ns, size=64, c1=32, c2=64, classes=10):
    """cfg4 — examples/4_mnist_res.py:38-44 topology (l1, l2=f(l1), Add, one more
    block, classifier) with convolutions substituted."""
    L = ns.L
    l_in = L.Input(shape=(3, size, size))
    l1 = L.Relu(L.BatchNorm(L.Conv2D(l_in, filters=c1, kernel_size=(3, 3), padding="same")))
    l = L.DepthwiseConv2D(l1, kernel_size=(3, 3), padding="same")
    l = L.PointwiseConv2D(l, filters=c1)
    l2 = L.Relu(L.BatchNorm(l))
    l = L.Add([l1, l2])
    l = L.Conv2D(l, filters=c2, kernel_size=(3, 3), strides=(2, 2), padding="same")
    l = L.Relu(L.BatchNorm(l))
    l = L.MaxPool2D(l, pool_size=(2, 2), strides=(2, 2))
    l = L.Flatten(l)
    l_out = L.Softmax(L.Dense(l, classes))
    return _build(ns, l_in, l_out, ns.opt.Adam(lr=0.001), ns.losses.CrossEntropy(),
                  ns.metrics.CategoricalAccuracy(), True)


def rnn_sentiment(ns, vocab=1000, emb=8, seq=256, hidden=32, trunc=4):
    """cfg5 — examples/9_rnn_sentiment.py:73-93 with the commented Embedding enabled."""
    L = ns.L
    l_in = L.Input(shape=(seq,))
    l = L.Embedding(l_in, input_dim=vocab, output_dim=emb, input_length=seq)
    l = L.SimpleRNN(l, hidden_dim=hidden, stateful=False, return_sequences=False, unroll=False, bptt_truncate=trunc)
    l = L.Dense(l, units=1)
    l_out = L.Sigmoid(l)
    return _build(ns, l_in, l_out, ns.opt.Adam(lr=0.01), ns.losses.BinaryCrossEntropy(),
                  ns.metrics.BinaryAccuracy(), True)


# name -> (builder, kwargs for the SMALL golden variant, input maker, target maker)
def make_data(name, n, rs, **kw):
    """Seeded synthetic data of the config's shape."""
    if name == "iris_mlp":
        x = rs.randn(n, kw.get("features", 4))
        y = np.eye(kw.get("classes", 3))[rs.randint(0, kw.get("classes", 3), n)]
    elif name == "mnist_mlp":
        x = rs.rand(n, 28, 28)
        y = np.eye(kw.get("classes", 10))[rs.randint(0, kw.get("classes", 10), n)]
    elif name == "mnist_cnn":
        s = kw.get("size", 28)
        x = rs.rand(n, 1, s, s)
        y = np.eye(kw.get("classes", 10))[rs.randint(0, kw.get("classes", 10), n)]
    elif name == "res_conv":
        s = kw.get("size", 64)
        x = rs.rand(n, 3, s, s)
        y = np.eye(kw.get("classes", 10))[rs.randint(0, kw.get("classes", 10), n)]
    elif name == "rnn_sentiment":
        x = rs.randint(0, kw.get("vocab", 1000), (n, kw.get("seq", 256)))
        y = rs.randint(0, 2, (n, 1)).astype(float)
    else:
        raise KeyError(name)
    return x, y


BUILDERS = dict(iris_mlp=iris_mlp, mnist_mlp=mnist_mlp, mnist_cnn=mnist_cnn, res_conv=res_conv,
                rnn_sentiment=rnn_sentiment)

# Small variants recorded in tests/golden/net_*.npz: (kwargs, batch, steps)
GOLDEN_NETS = {
    "iris_mlp": (dict(), 24, 6),
    "mnist_mlp": (dict(h1=16, h2=8), 8, 4),
    "mnist_cnn": (dict(f1=2, f2=4), 6, 4),
    "res_conv": (dict(size=8, c1=4, c2=8), 4, 3),
    "rnn_sentiment": (dict(vocab=50, emb=3, seq=7, hidden=5, trunc=2), 3, 3),
}
