"""NumPy float64 restatement of the dnpy hot path — TEST INFRASTRUCTURE ONLY.

Every function cites the reference file:line it restates (paths relative to
``/root/reference``).  The arithmetic is the reference's *literal* arithmetic,
quirks included; only the loop structure differs (vectorised
over output positions where the reference loops in Python).  The ``*_loops``
variants keep the reference's loop structure (one NumPy call per filter/output
position) and are what ``bench.py`` times as the CPU baseline, because that is
what the reference's CPU path costs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import this module.  Parity is pinned by ``tests/test_oracle.py``
against fixtures recorded from the real reference (tests/golden/make_golden.py).
"""
import numpy as np
from numpy.lib.stride_tricks import sliding_window_view

EPS = 10e-8  # dnpy/layers/base.py:22, dnpy/losses.py:11 (note: 1e-7)


# ----------------------------------------------------------------------------
# Shape helpers — dnpy/utils.py:52-109
# ----------------------------------------------------------------------------
def get_output(input_size, kernel_size, strides, padding, dilation_rate=None):
    """dnpy/utils.py:91-109."""
    i = np.array(input_size)
    k = np.array(kernel_size)
    s = np.array(strides)
    d = np.array(dilation_rate) if dilation_rate else np.ones_like(k)
    if isinstance(padding, str) and padding in ("none", "valid"):
        o = np.ceil((i - d * (k - 1)) / s)
    elif isinstance(padding, str) and padding in ("same", "zeros"):
        o = np.ceil(i / s)
    elif isinstance(padding, (list, tuple, np.ndarray)):
        o = np.floor((i - k + 2 * np.array(padding)) / s) + 1
    else:
        raise ValueError("Unknown padding")
    return o.astype(int)


def get_padding(padding, input_size, output_size, kernel_size, strides):
    """dnpy/utils.py:52-76 (total pad per axis)."""
    i = np.array(input_size)
    o = np.array(output_size)
    k = np.array(kernel_size)
    s = np.array(strides)
    mode = str(padding).lower()
    if mode in ("none", "valid"):
        p = np.zeros_like(k)
    elif mode in ("same", "zeros"):
        p = np.ceil(s * (o - 1) - i + k)
    else:
        raise ValueError("Unknown padding")
    return p.astype(int)


def get_side_paddings(pads):
    """dnpy/utils.py:79-88: floor half before, the rest after."""
    return tuple((int(p) // 2, int(p) - int(p) // 2) for p in pads)


def _pad2d(x, pt, pb, pl, pr):
    """Zero-padded copy of an NCHW batch (convolution.py:203-207, pooling.py:53-57)."""
    b, c, h, w = x.shape
    xp = np.zeros((b, c, h + pt + pb, w + pl + pr))
    xp[:, :, pt:pt + h, pl:pl + w] = x
    return xp


def _crop2d(xp, pt, pb, pl, pr):
    h, w = xp.shape[2], xp.shape[3]
    return xp[:, :, pt:h - pb, pl:w - pr]


def _windows(xp, ky, kx, sy, sx, oy, ox):
    """(B,C,OY,OX,ky,kx) strided window view of the padded input."""
    win = sliding_window_view(xp, (ky, kx), axis=(2, 3))
    return win[:, :, ::sy, ::sx][:, :, :oy, :ox]


# ----------------------------------------------------------------------------
# Regularizers — dnpy/regularizers.py:25-26,38-39,52-53
# ----------------------------------------------------------------------------
def reg_backward(param, l1=0.0, l2=0.0):
    """lambda_l1*sign(p) + lambda_l2*2p; (0,0) means "no regularizer"."""
    out = np.zeros_like(param, dtype=float)
    if l1:
        out = out + l1 * np.sign(param)
    if l2:
        out = out + l2 * (2 * param)
    return out


# ----------------------------------------------------------------------------
# Dense — dnpy/layers/core.py:116-135
# ----------------------------------------------------------------------------
def dense_forward(x, w, b):
    return np.dot(x, w) + b


def dense_backward(x, w, b, delta, reg_w=(0.0, 0.0), reg_b=(0.0, 0.0)):
    """Returns (dx, gw_inc, gb_inc); regulariser sits inside the /m (Q11)."""
    m = delta.shape[0]
    dx = np.dot(delta, w.T)
    gw = np.dot(x.T, delta)
    gb = np.sum(delta, axis=0, keepdims=True)
    if any(reg_w):
        gw = gw + reg_backward(w, *reg_w)
    if any(reg_b):
        gb = gb + reg_backward(b, *reg_b)
    return dx, gw / m, gb / m


# ----------------------------------------------------------------------------
# Conv2D / PointwiseConv2D — dnpy/layers/convolution.py:193-273 (live definition)
# ----------------------------------------------------------------------------
def conv2d_forward(x, w, b, strides, pads, oshape):
    """pads = (top, bottom, left, right); oshape = (F, OY, OX).

    Bias enters inside the window sum, i.e. K*b with K=C*ky*kx (Q1,
    convolution.py:223-224).
    """
    f, c, ky, kx = w.shape
    sy, sx = strides
    _, oy, ox = oshape
    xp = _pad2d(x, *pads)
    win = _windows(xp, ky, kx, sy, sx, oy, ox)
    out = np.einsum('bcyxij,fcij->bfyx', win, w, optimize=True)
    return out + (c * ky * kx) * b.reshape(1, f, 1, 1)


def conv2d_backward(x, w, b, delta, strides, pads, reg_w=(0.0, 0.0), reg_b=(0.0, 0.0)):
    """Returns (dx, gw_inc, gb_inc) — convolution.py:227-273.

    gw/gb are batch MEANS per output position summed over positions; the
    regulariser term is added once per output position (Q2).
    """
    bsz = x.shape[0]
    f, c, ky, kx = w.shape
    sy, sx = strides
    oy, ox = delta.shape[2], delta.shape[3]
    xp = _pad2d(x, *pads)
    dxp = np.zeros_like(xp)
    gw = np.zeros_like(w, dtype=float)
    for i in range(ky):
        for j in range(kx):
            sl = (slice(None), slice(None), slice(i, i + sy * oy, sy), slice(j, j + sx * ox, sx))
            dxp[sl] += np.einsum('bfyx,fc->bcyx', delta, w[:, :, i, j], optimize=True)
            gw[:, :, i, j] = np.einsum('bfyx,bcyx->fc', delta, xp[sl], optimize=True) / bsz
    gb = delta.sum(axis=(0, 2, 3)) / bsz
    if any(reg_w):
        gw = gw + (oy * ox) * reg_backward(w, *reg_w)
    if any(reg_b):
        gb = gb + (oy * ox) * reg_backward(b, *reg_b)
    return _crop2d(dxp, *pads), gw, gb


def conv2d_forward_loops(x, w, b, strides, pads, oshape):
    """Same arithmetic with the reference's (filter, oy, ox) loop structure
    (convolution.py:209-225) — the CPU-baseline cost model."""
    f, c, ky, kx = w.shape
    sy, sx = strides
    _, oy, ox = oshape
    xp = _pad2d(x, *pads)
    out = np.zeros((x.shape[0], f, oy, ox))
    for fi in range(f):
        for y in range(oy):
            for xx in range(ox):
                patch = xp[:, :, y * sy:y * sy + ky, xx * sx:xx * sx + kx]
                out[:, fi, y, xx] = np.sum(patch * w[fi] + b[fi], axis=(1, 2, 3))
    return out


def conv2d_backward_loops(x, w, b, delta, strides, pads, reg_w=(0.0, 0.0), reg_b=(0.0, 0.0)):
    """Reference loop structure of convolution.py:236-273."""
    f, c, ky, kx = w.shape
    sy, sx = strides
    oy, ox = delta.shape[2], delta.shape[3]
    xp = _pad2d(x, *pads)
    dxp = np.zeros_like(xp)
    gw = np.zeros_like(w, dtype=float)
    gb = np.zeros(f)
    for fi in range(f):
        for y in range(oy):
            for xx in range(ox):
                d = delta[:, fi, y, xx]
                sl = (slice(None), slice(None), slice(y * sy, y * sy + ky), slice(xx * sx, xx * sx + kx))
                dxp[sl] += np.outer(d, w[fi]).reshape((len(d),) + w[fi].shape)
                g_w = np.mean(d[:, None, None, None] * xp[sl], axis=0)
                g_b = np.mean(d, axis=0)
                if any(reg_w):
                    g_w = g_w + reg_backward(w[fi], *reg_w)
                if any(reg_b):
                    g_b = g_b + reg_backward(b[fi], *reg_b)
                gw[fi] += g_w
                gb[fi] += g_b
    return _crop2d(dxp, *pads), gw, gb


# ----------------------------------------------------------------------------
# DepthwiseConv2D — dnpy/layers/convolution.py:323-388
# ----------------------------------------------------------------------------
def dwconv2d_forward(x, w, b, strides, pads, oshape):
    """w is (C,1,ky,kx); bias enters as (ky*kx)*b (convolution.py:344-347)."""
    c, _, ky, kx = w.shape
    sy, sx = strides
    _, oy, ox = oshape
    xp = _pad2d(x, *pads)
    win = _windows(xp, ky, kx, sy, sx, oy, ox)
    out = np.einsum('bcyxij,cij->bcyx', win, w[:, 0], optimize=True)
    return out + (ky * kx) * b.reshape(1, c, 1, 1)


def dwconv2d_backward(x, w, b, delta, strides, pads, reg_w=(0.0, 0.0), reg_b=(0.0, 0.0)):
    """convolution.py:350-388.  A kernel regularizer makes the reference raise
    (``g_w1`` is (ky,kx) but the term is (1,ky,kx), convolution.py:380), so it
    is rejected here as well."""
    if any(reg_w):
        raise ValueError("DepthwiseConv2D with a kernel_regularizer raises in the reference (convolution.py:380)")
    bsz = x.shape[0]
    c, _, ky, kx = w.shape
    sy, sx = strides
    oy, ox = delta.shape[2], delta.shape[3]
    xp = _pad2d(x, *pads)
    dxp = np.zeros_like(xp)
    gw = np.zeros_like(w, dtype=float)
    for i in range(ky):
        for j in range(kx):
            sl = (slice(None), slice(None), slice(i, i + sy * oy, sy), slice(j, j + sx * ox, sx))
            dxp[sl] += delta * w[:, 0, i, j].reshape(1, c, 1, 1)
            gw[:, 0, i, j] = np.einsum('bcyx,bcyx->c', delta, xp[sl]) / bsz
    gb = delta.sum(axis=(0, 2, 3)) / bsz
    if any(reg_b):
        gb = gb + (oy * ox) * reg_backward(b, *reg_b)
    return _crop2d(dxp, *pads), gw, gb


# ----------------------------------------------------------------------------
# MaxPool2D / AvgPool2D — dnpy/layers/pooling.py:43-99, 118-174
# ----------------------------------------------------------------------------
def maxpool2d_forward(x, pool, strides, pads, oshape):
    """Zero padding takes part in the max (Q6); first max wins (np.argmax over
    the row-major flattened window, pooling.py:74-76).  Returns (out, idx)."""
    py, px = pool
    sy, sx = strides
    _, oy, ox = oshape
    xp = _pad2d(x, *pads)
    win = _windows(xp, py, px, sy, sx, oy, ox).reshape(x.shape[0], x.shape[1], oy, ox, py * px)
    return win.max(axis=-1), win.argmax(axis=-1).astype(np.int64)


def maxpool2d_backward(delta, idx, in_shape, pool, strides, pads, batch_route="literal"):
    """pooling.py:78-99.

    ``literal`` reproduces pooling.py:95 exactly (Q5): the flat index into the
    (B,py,px) window view is the *per-sample* offset in [0,py*px), so every
    sample's delta lands in sample 0's window, and NumPy's fancy-index ``+=``
    (gather, add, scatter) makes the LAST sample that selected an offset win.
    ``per_sample`` is the intended scatter-add (what the literal code does when
    B == 1, the only regime the reference's own tests cover).
    """
    py, px = pool
    sy, sx = strides
    b, c, oy, ox = delta.shape
    _, _, h, w = in_shape
    pt, pb, pl, pr = pads
    dxp = np.zeros((b, c, h + pt + pb, w + pl + pr))
    barange = np.arange(b)[:, None]
    for y in range(oy):
        for x in range(ox):
            sel = idx[:, :, y, x]                      # (B,C) offsets in [0,P)
            d = delta[:, :, y, x]
            for p in range(py * px):
                iy, ix = y * sy + p // px, x * sx + p % px
                hit = (sel == p)
                if batch_route == "per_sample":
                    dxp[:, :, iy, ix] += d * hit
                else:
                    last = np.where(hit, barange, -1).max(axis=0)      # (C,) last sample choosing p
                    has = last >= 0
                    vals = d[np.clip(last, 0, None), np.arange(c)]
                    dxp[0, :, iy, ix] += np.where(has, vals, 0.0)
    return _crop2d(dxp, *pads)


def maxpool2d_backward_loops(delta, idx, in_shape, pool, strides, pads):
    """The literal statement itself (pooling.py:86-99) — used to pin the
    vectorised ``literal`` route above."""
    py, px = pool
    sy, sx = strides
    b, c, oy, ox = delta.shape
    _, _, h, w = in_shape
    pt, pb, pl, pr = pads
    dxp = np.zeros((b, c, h + pt + pb, w + pl + pr))
    for z in range(c):
        for y in range(oy):
            for x in range(ox):
                dxp[:, z, y * sy:y * sy + py, x * sx:x * sx + px].flat[idx[:, z, y, x]] += delta[:, z, y, x]
    return _crop2d(dxp, *pads)


def avgpool2d_forward(x, pool, strides, pads, oshape):
    """Mean over the zero-padded window, always /(py*px) (pooling.py:118-147)."""
    py, px = pool
    sy, sx = strides
    _, oy, ox = oshape
    xp = _pad2d(x, *pads)
    win = _windows(xp, py, px, sy, sx, oy, ox)
    return win.mean(axis=(-1, -2))


def avgpool2d_backward(delta, in_shape, pool, strides, pads):
    """pooling.py:149-174: the window is ASSIGNED (later windows overwrite) and
    the tile count 4 is hard-coded, so only 2x2 pools are valid."""
    py, px = pool
    if py * px != 4:
        raise ValueError("AvgPool2D.backward only works for 4-element pools (pooling.py:169)")
    sy, sx = strides
    b, c, oy, ox = delta.shape
    _, _, h, w = in_shape
    pt, pb, pl, pr = pads
    dxp = np.zeros((b, c, h + pt + pb, w + pl + pr))
    for y in range(oy):
        for x in range(ox):
            dxp[:, :, y * sy:y * sy + py, x * sx:x * sx + px] = (delta[:, :, y, x] * 1 / (py * px))[:, :, None, None]
    return _crop2d(dxp, *pads)


# ----------------------------------------------------------------------------
# Activations — dnpy/layers/activations.py
# ----------------------------------------------------------------------------
def leakyrelu_forward(x, alpha):
    """activations.py:17-19 (Relu is alpha=0, :25-28)."""
    return np.where(x > 0, x, x * alpha)


def leakyrelu_backward(x, delta, alpha):
    """activations.py:21-22 (gate = x>0 cached in forward)."""
    return np.where(x > 0, delta, delta * alpha)


def prelu_forward(x, alpha):
    """activations.py:59-61; alpha has the layer's full oshape."""
    return np.where(x > 0, x, x * alpha)


def prelu_backward(x, delta, alpha):
    """activations.py:63-66 -> (dx, galpha_inc)."""
    gate = x > 0
    dx = np.where(gate, delta, delta * alpha)
    return dx, np.mean(delta * np.invert(gate).astype(float), axis=0)


def sigmoid_forward(x):
    """activations.py:77-78 (epsilon inside and outside the exp, Q18)."""
    return 1.0 / (1.0 + np.exp(-x + EPS) + EPS)


def sigmoid_backward(y, delta):
    """activations.py:80-81."""
    return delta * (y * (1 - y))


def tanh_forward(x):
    """activations.py:92-95 (naive exp formula)."""
    a, b = np.exp(x), np.exp(-x)
    return (a - b) / (a + b)


def tanh_backward(y, delta):
    """activations.py:97-98."""
    return delta * (1 - y ** 2)


def softmax_forward(x, stable=True):
    """activations.py:116-124."""
    z = x - np.max(x, axis=1, keepdims=True) if stable else x
    e = np.exp(z)
    return e / np.sum(e, axis=1, keepdims=True)


def softmax_backward(y, delta, ce_loss=False):
    """activations.py:126-135: pass-through for the CE shortcut, else the full
    per-row Jacobian product delta·(diag(s) − s sᵀ)."""
    if ce_loss:
        return delta
    return y * (delta - np.sum(delta * y, axis=1, keepdims=True))


def logsoftmax_forward(x, stable=True):
    """activations.py:152-160."""
    z = x - np.max(x, axis=1, keepdims=True) if stable else x
    return z - np.log(np.sum(np.exp(z), axis=1, keepdims=True) + EPS)


# ----------------------------------------------------------------------------
# Dropout / BatchNorm / Add — regularization.py:17-25, 94-156; merge.py:13-20
# ----------------------------------------------------------------------------
def dropout_mask(shape, rate, rng=np.random):
    """regularization.py:19: one draw from the global stream per forward."""
    return (rng.random(shape) > rate).astype(float)


def dropout_forward(x, rate, training, gate=None):
    """Non-inverted dropout (Q12)."""
    return x * gate if training else x * (1 - rate)


def dropout_backward(delta, gate):
    return delta * gate


def gaussian_noise_forward(x, mean, stddev, training, rng=np.random):
    """regularization.py:39-44: one ``normal(mean, stddev, x.shape)`` draw from the global stream per TRAINING forward;
    identity in test mode.  Backward is a pass-through copy (regularization.py:46-47)."""
    if not training:
        return x
    return x + rng.normal(loc=mean, scale=stddev, size=np.shape(x))


def batchnorm_forward(x, gamma, beta, moving_mu, moving_var, momentum, training):
    """regularization.py:94-136 with bias_correction=False.

    Statistics over axis 0 only (Q8).  Returns (out, cache, new_mu, new_var)."""
    if training:
        mu = np.mean(x, axis=0, keepdims=True)
        var = np.var(x, axis=0, keepdims=True)
        new_mu = momentum * moving_mu + (1.0 - momentum) * mu
        new_var = momentum * moving_var + (1.0 - momentum) * var
    else:
        mu, var = moving_mu, moving_var
        new_mu, new_var = moving_mu, moving_var
    std = np.sqrt(var + EPS)
    xn = (x - mu) / std
    return gamma * xn + beta, dict(mu=mu, var=var, std=std, xn=xn), new_mu, new_var


def batchnorm_backward(delta, gamma, cache):
    """regularization.py:138-156, literal (Q7): dgamma uses mu, dX multiplies
    by std.  Returns (dx, ggamma_inc, gbeta_inc)."""
    m = delta.shape[0]
    dxn = delta * gamma
    dx = (1.0 / m) * cache['std'] * (m * dxn - np.sum(dxn, axis=0, keepdims=True)
                                     - cache['xn'] * np.sum(dxn * cache['xn'], axis=0, keepdims=True))
    return dx, np.sum(delta * cache['mu'], axis=0), np.sum(delta, axis=0)


# ----------------------------------------------------------------------------
# Embedding — dnpy/layers/core.py:54-65
# ----------------------------------------------------------------------------
def embedding_forward(idx, w):
    return w[idx]


def embedding_backward_inplace(gw, idx, delta):
    """core.py:62-65: batch-mean first, then NumPy fancy-index ``+=`` (gather,
    add, scatter: the last (b,l) in row-major order wins, Q15)."""
    gw[idx] += np.mean(delta, axis=0)


# ----------------------------------------------------------------------------
# SimpleRNN — dnpy/layers/recurrent.py:12-23, 153-266 (non-stateful, last output)
# ----------------------------------------------------------------------------
def rnn_forward(x, waa, wax, ba):
    """x (B,T,D) -> states (B,T,H); h0=0; output is states[:, -1]."""
    b, t, _ = x.shape
    hdim = waa.shape[0]
    h = np.zeros((b, hdim))
    states = np.zeros((b, t, hdim))
    for ti in range(t):
        h = np.tanh(np.dot(h, waa.T) + np.dot(x[:, ti], wax.T) + ba)
        states[:, ti] = h
    return states


def rnn_backward(x, states, delta, waa, wax, bptt_truncate=None):
    """recurrent.py:190-266, vectorised over the batch (samples are
    independent; gradients are batch SUMS, Q16/Q17).

    Returns (dx (B,T,D), gwaa_inc, gwax_inc, gba_inc)."""
    b, t, d = x.shape
    hdim = waa.shape[0]
    trunc = bptt_truncate if bptt_truncate else t
    gwaa = np.zeros_like(waa, dtype=float)
    gwax = np.zeros_like(wax, dtype=float)
    gba = np.zeros((1, hdim))
    dx = np.zeros((b, t, d))
    dh_prev = np.zeros((b, hdim))
    zeros_h = np.zeros((b, hdim))
    for ti in reversed(range(t)):
        h_t = states[:, ti]
        dy = delta if ti == t - 1 else zeros_h
        pre = (dy + dh_prev) * (1 - h_t ** 2)
        s = np.array(pre)
        for k in reversed(range(max(0, ti - trunc), ti + 1)):
            h_km1 = states[:, k - 1] if k > 0 else zeros_h
            gwaa += np.dot(s.T, h_km1)
            gwax += np.dot(s.T, x[:, k])
            gba += s.sum(axis=0, keepdims=True)
            s = np.dot(s * (1 - h_km1 ** 2), waa)
        dh_prev = np.dot(pre, waa)
        dx[:, ti] = np.dot(pre, wax)
    return dx, gwaa, gwax, gba


# ----------------------------------------------------------------------------
# Losses — dnpy/losses.py:17-26, 60-94
# ----------------------------------------------------------------------------
def mse_loss(p, t):
    """losses.py:22-23: float() of the per-column mean => single-column outputs only."""
    return float(np.mean((p - t) ** 2, axis=0, keepdims=True).item())


def mse_delta(p, t):
    return 2 * (p - t)


def rmse_loss(p, t):
    """losses.py:34-37."""
    return float(np.sqrt(np.mean((p - t) ** 2, axis=0, keepdims=True)).item())


def rmse_delta(p, t):
    """losses.py:39-42."""
    return (p - t) / np.sqrt((p - t) ** 2 + EPS)


def mae_loss(p, t):
    """losses.py:50-53."""
    return float(np.mean(np.abs(p - t), axis=0, keepdims=True).item())


def mae_delta(p, t):
    return np.sign(p - t)


def nll_loss(p, t):
    """losses.py:105-110 (p holds log-probabilities)."""
    v = np.sum(t.astype(float) * p, axis=1, keepdims=True)
    return -1.0 * float(np.mean(v, axis=0, keepdims=True).item())


def nll_delta(p, t):
    return np.exp(p) - t


def hinge_loss(p, t):
    """losses.py:121-127."""
    return float(np.mean(np.maximum(0, (1 - t * p)), axis=0, keepdims=True).item())


def hinge_delta(p, t):
    """losses.py:129-133."""
    gate = np.invert((t * p) > 1).astype(float)
    return gate * (-1.0 * t)


def bce_loss(p, t):
    """losses.py:65-68 (single output column)."""
    v = t * np.log(p + EPS) + (1.0 - t) * np.log(1.0 - p + EPS)
    return -1.0 * float(np.mean(v, axis=0, keepdims=True).reshape(-1)[0])


def bce_delta(p, t):
    """losses.py:70-73."""
    return -1.0 * (t * 1.0 / (p + EPS) + (1.0 - t) * 1 / (1.0 - p + EPS) * -1.0)


def ce_loss(p, t):
    """losses.py:82-86."""
    v = np.sum(t.astype(float) * np.log(p + EPS), axis=1, keepdims=True)
    return -1.0 * float(np.mean(v, axis=0, keepdims=True).reshape(-1)[0])


def ce_delta(p, t, softmax_output=False):
    """losses.py:88-94; never divided by the batch size (Q17)."""
    if softmax_output:
        return p - t
    return -1.0 * (t.astype(float) * 1 / (p + EPS))


# ----------------------------------------------------------------------------
# Optimizers — dnpy/optimizers.py:31-55, 73-84, 106-126 (in place on p / state)
# ----------------------------------------------------------------------------
def sgd_apply(p, g, v, lr, momentum=0.0, nesterov=False):
    """Dampened EMA momentum (Q13)."""
    v_prev = v.copy()
    v[...] = momentum * v + (1 - momentum) * g
    new_v = v if not nesterov else -momentum * v_prev + (1 + momentum) * v
    p -= lr * new_v


def rmsprop_apply(p, g, s, lr, rho=0.9, eps=EPS):
    s[...] = rho * s + (1.0 - rho) * g ** 2
    p -= lr * (g / np.sqrt(s + eps))


def adam_apply(p, g, v, s, lr, beta_1=0.9, beta_2=0.999, eps=EPS):
    """No bias correction, epsilon under the sqrt (Q13)."""
    v[...] = beta_1 * v + (1.0 - beta_1) * g
    s[...] = beta_2 * s + (1.0 - beta_2) * g ** 2
    p -= lr * (v / np.sqrt(s + eps))
