"""Per-operator differential cases shared by the golden generator and the tests.

Each case drives ONE layer exactly the way the reference's own unit tests do
(tests/coverage/test_layers.py:71-82): ``l0 = Input(...); l0.output = x;
l1 = Layer(l0, ...); l1.initialize(); l1.forward(); l1.delta = d; l1.backward()``.
Inputs are seeded; every parameter (bias included) is overwritten with seeded
non-zero values so the quirks the reference's own goldens leave unpinned
 are pinned here.
"""
import numpy as np

# kind, input shape(s), ctor kwargs (regularizers/initializers encoded as lists), batch, training
LAYER_CASES = [
    dict(name="dense_reg", kind="Dense", in_shape=(7,), batch=6,
         kw=dict(units=5, kernel_regularizer=["L2", 0.01], bias_regularizer=["L1", 0.02])),
    dict(name="dense_plain", kind="Dense", in_shape=(9,), batch=1, kw=dict(units=3)),
    dict(name="conv_3x3_same", kind="Conv2D", in_shape=(3, 7, 6), batch=3,
         kw=dict(filters=4, kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="conv_2x2_s2_none", kind="Conv2D", in_shape=(2, 8, 8), batch=2,
         kw=dict(filters=3, kernel_size=(2, 2), strides=(2, 2), padding="none")),
    dict(name="conv_3x3_s2_same_reg", kind="Conv2D", in_shape=(3, 9, 9), batch=4,
         kw=dict(filters=2, kernel_size=(3, 3), strides=(2, 2), padding="same",
                 kernel_regularizer=["L1L2", 0.01, 0.02], bias_regularizer=["L2", 0.03])),
    dict(name="conv_3x3_none_c1", kind="Conv2D", in_shape=(1, 10, 10), batch=5,
         kw=dict(filters=2, kernel_size=(3, 3), strides=(1, 1), padding="none")),
    dict(name="conv_5x3_same", kind="Conv2D", in_shape=(2, 8, 7), batch=2,
         kw=dict(filters=3, kernel_size=(5, 3), strides=(1, 1), padding="same")),
    dict(name="pointwise", kind="PointwiseConv2D", in_shape=(4, 5, 5), batch=3, kw=dict(filters=3)),
    dict(name="depthwise_3x3_same", kind="DepthwiseConv2D", in_shape=(3, 6, 6), batch=3,
         kw=dict(kernel_size=(3, 3), strides=(1, 1), padding="same")),
    dict(name="depthwise_2x2_s2_reg", kind="DepthwiseConv2D", in_shape=(2, 6, 6), batch=2,
         kw=dict(kernel_size=(2, 2), strides=(2, 2), padding="none",
                 bias_regularizer=["L1", 0.01])),   # kernel_regularizer raises in the reference (shape bug, convolution.py:380)
    dict(name="maxpool_3x3_s2_b4", kind="MaxPool2D", in_shape=(3, 7, 7), batch=4,
         kw=dict(pool_size=(3, 3), strides=(2, 2), padding="none")),
    dict(name="maxpool_2x2_same_neg", kind="MaxPool2D", in_shape=(2, 5, 5), batch=3, negative=True,
         kw=dict(pool_size=(2, 2), strides=(2, 2), padding="same")),
    dict(name="maxpool_2x2_b1", kind="MaxPool2D", in_shape=(2, 6, 6), batch=1,
         kw=dict(pool_size=(2, 2), strides=(2, 2), padding="none")),
    dict(name="maxpool_ties", kind="MaxPool2D", in_shape=(2, 6, 6), batch=3, quantize=True,
         kw=dict(pool_size=(3, 3), strides=(1, 1), padding="none")),
    dict(name="globalmaxpool", kind="GlobalMaxPool2D", in_shape=(3, 4, 4), batch=2, kw=dict()),
    dict(name="avgpool_2x2", kind="AvgPool2D", in_shape=(2, 6, 6), batch=3,
         kw=dict(pool_size=(2, 2), strides=(2, 2), padding="none")),
    dict(name="avgpool_2x2_same", kind="AvgPool2D", in_shape=(2, 5, 5), batch=2,
         kw=dict(pool_size=(2, 2), strides=(2, 2), padding="same")),
    # pooling.py:177-181: the window is the whole map; the reference's backward only exists for 4-element windows
    dict(name="globalavgpool_2x2", kind="GlobalAvgPool2D", in_shape=(3, 2, 2), batch=3, kw=dict()),
    dict(name="globalavgpool_fwd", kind="GlobalAvgPool2D", in_shape=(2, 5, 4), batch=3, kw=dict(), backward=False),
    dict(name="leakyrelu", kind="LeakyRelu", in_shape=(4, 3), batch=5, kw=dict(alpha=0.3)),
    dict(name="relu", kind="Relu", in_shape=(11,), batch=4, kw=dict()),
    dict(name="prelu", kind="PRelu", in_shape=(6,), batch=5, kw=dict()),
    dict(name="sigmoid", kind="Sigmoid", in_shape=(5,), batch=4, kw=dict()),
    dict(name="tanh", kind="Tanh", in_shape=(5,), batch=4, kw=dict()),
    dict(name="softmax_generic", kind="Softmax", in_shape=(6,), batch=5, kw=dict()),
    dict(name="logsoftmax", kind="LogSoftmax", in_shape=(6,), batch=5, kw=dict()),
    dict(name="dropout_train", kind="Dropout", in_shape=(8,), batch=6, kw=dict(rate=0.3), training=True, seed=11),
    dict(name="dropout_test", kind="Dropout", in_shape=(8,), batch=6, kw=dict(rate=0.3), training=False,
         backward=False),
    # regularization.py:39-47: one np.random.normal call per training forward, from the global stream
    dict(name="gaussiannoise_train", kind="GaussianNoise", in_shape=(3, 4), batch=5, kw=dict(mean=0.1, stddev=0.5),
         training=True, seed=13),
    dict(name="gaussiannoise_test", kind="GaussianNoise", in_shape=(7,), batch=4, kw=dict(mean=0.1, stddev=0.5),
         training=False),
    dict(name="batchnorm_train_1d", kind="BatchNorm", in_shape=(5,), batch=7, kw=dict(), training=True),
    dict(name="batchnorm_train_3d", kind="BatchNorm", in_shape=(2, 3, 3), batch=4, kw=dict(momentum=0.9), training=True),
    dict(name="batchnorm_test", kind="BatchNorm", in_shape=(5,), batch=3, kw=dict(), training=False),
    dict(name="add3", kind="Add", in_shape=(4,), n_inputs=3, batch=3, kw=dict()),
    dict(name="subtract3", kind="Subtract", in_shape=(5,), n_inputs=3, batch=4, kw=dict()),
    dict(name="multiply3", kind="Multiply", in_shape=(2, 3), n_inputs=3, batch=3, kw=dict()),
    dict(name="divide2", kind="Divide", in_shape=(6,), n_inputs=2, batch=3, kw=dict()),
    dict(name="power2", kind="Power", in_shape=(6,), n_inputs=2, batch=3, positive=True, kw=dict()),
    dict(name="maximum3", kind="Maximum", in_shape=(7,), n_inputs=3, batch=4, quantize=True, kw=dict()),
    dict(name="minimum2", kind="Minimum", in_shape=(7,), n_inputs=2, batch=4, kw=dict()),
    dict(name="average3", kind="Average", in_shape=(3, 2), n_inputs=3, batch=5, kw=dict()),
    dict(name="embedding_dups", kind="Embedding", in_shape=(5,), batch=4, int_input=7,
         kw=dict(input_dim=7, output_dim=3, input_length=5)),
    dict(name="rnn_trunc2", kind="SimpleRNN", in_shape=(6, 3), batch=3, kw=dict(hidden_dim=4, bptt_truncate=2)),
    dict(name="rnn_full", kind="SimpleRNN", in_shape=(4, 2), batch=2, kw=dict(hidden_dim=3)),
]

LOSS_CASES = [
    dict(name="ce_generic", kind="CrossEntropy", shape=(6, 4), softmax_output=False),
    dict(name="ce_softmax", kind="CrossEntropy", shape=(6, 4), softmax_output=True),
    dict(name="bce", kind="BinaryCrossEntropy", shape=(7, 1)),
    dict(name="mse", kind="MSE", shape=(5, 1)),
    dict(name="rmse", kind="RMSE", shape=(6, 1)),
    dict(name="mae", kind="MAE", shape=(6, 1)),
    dict(name="nll", kind="NLL", shape=(5, 4)),
    dict(name="hinge", kind="Hinge", shape=(9, 1)),
]

OPT_CASES = [
    dict(name="adam", kind="Adam", kw=dict(lr=0.01)),
    dict(name="rmsprop", kind="RMSProp", kw=dict(lr=0.01, rho=0.8)),
    dict(name="sgd_mom", kind="SGD", kw=dict(lr=0.1, momentum=0.9)),
    dict(name="sgd_plain", kind="SGD", kw=dict(lr=0.1)),
    dict(name="sgd_nesterov", kind="SGD", kw=dict(lr=0.1, momentum=0.5, nesterov=True)),
]


def decode_kwargs(ns, kw):
    """Turn ["L2", 0.01]-style entries into regularizer objects of namespace ``ns``."""
    out = {}
    for k, v in kw.items():
        if k.endswith("_regularizer") and v is not None:
            out[k] = getattr(ns.reg, v[0])(*v[1:])
        else:
            out[k] = v
    return out


def case_inputs(case):
    """Seeded inputs/delta/param-overrides for a layer case (needs the output
    shape, so params/delta are drawn by ``draw_like`` once the layer exists)."""
    rs = np.random.RandomState(sum(map(ord, case["name"])))
    n_in = case.get("n_inputs", 1)
    xs = []
    for _ in range(n_in):
        if "int_input" in case:
            xs.append(rs.randint(0, case["int_input"], (case["batch"],) + tuple(case["in_shape"])))
        else:
            x = rs.randn(case["batch"], *case["in_shape"])
            if case.get("negative"):
                x = -np.abs(x) - 0.1
            if case.get("positive"):
                x = np.abs(x) + 0.25             # np.power on negative bases is NaN: keep the case meaningful
            if case.get("quantize"):
                x = np.round(x * 2) / 2      # many exact ties -> first-max-wins matters
            xs.append(x)
    return rs, xs


def draw_params(rs, params, scale=0.5):
    """Seeded non-zero values for every parameter, in sorted key order."""
    out = {}
    for k in sorted(params.keys()):
        v = rs.randn(*np.shape(params[k])) * scale
        if k == "moving_var":
            v = np.abs(v) + 0.5
        out[k] = v
    return out
