"""Host-side helpers with the reference's exact semantics (dnpy/utils.py).

Plotting helpers (utils.py:35-49) are out of scope; the shape helpers are the
ones the layer constructors depend on and follow utils.py:52-109 to the letter,
including "first the output size, then the padding" and the floor-half-before /
rest-after split of "same" padding.
"""
import copy
import random

import numpy as np


def to_categorical(y, num_classes):
    """utils.py:8-11."""
    out = np.zeros((len(y), num_classes))
    out[np.arange(len(y)), y] = 1.0
    return out


def to_binary(y, label):
    """utils.py:14-17."""
    out = np.zeros_like(y)
    out[y == label] = 1.0
    return out


def to_label(y):
    return np.argmax(y, axis=1)


def shuffle_dataset(x, y):
    """utils.py:24-27 (consumes the global NumPy stream exactly like the reference)."""
    order = np.arange(len(x))
    np.random.shuffle(order)
    return x[order], y[order]


def split_dataset(x, y, split=0.8):
    n = int(len(x) * split)
    return x[:n], y[:n]


def get_output(input_size, kernel_size, strides, padding, dilation_rate=None):
    """Output extent per axis (utils.py:91-109)."""
    isz, ksz, st = np.array(input_size), np.array(kernel_size), np.array(strides)
    dil = np.array(dilation_rate) if dilation_rate else np.ones_like(ksz)
    assert isz.ndim == ksz.ndim == st.ndim == dil.ndim
    if isinstance(padding, str) and padding in {"none", "valid"}:
        out = np.ceil((isz - dil * (ksz - 1)) / st)
    elif isinstance(padding, str) and padding in {"same", "zeros"}:
        out = np.ceil(isz / st)
    elif isinstance(padding, (list, tuple, np.ndarray)):
        out = np.floor((isz - ksz + 2 * np.array(padding)) / st) + 1
    else:
        raise ValueError("Unknown padding")
    return out.astype(int)


def get_padding(padding, input_size, output_size, kernel_size, strides):
    """Total zero padding per axis (utils.py:52-76)."""
    isz, osz = np.array(input_size), np.array(output_size)
    ksz, st = np.array(kernel_size), np.array(strides)
    mode = str(padding).lower()
    if mode in {"none", "valid"}:
        pads = np.zeros_like(ksz)
    elif mode in {"same", "zeros"}:
        pads = np.ceil(st * (osz - 1) - isz + ksz)
    else:
        raise ValueError("Unknown padding")
    return pads.astype(int)


def get_side_paddings(pads):
    """((before, after), ...) with the larger half after (utils.py:79-88)."""
    return tuple((p // 2, p - p // 2) for p in pads)


def params2vector(params):
    """Flatten a list of param dicts (utils.py:112-124) — used by gradient checking."""
    chunks, slices, pos = [], [], 0
    for layer_params in params:
        for _, v in layer_params.items():
            v = np.asarray(v)
            chunks.append(v.reshape(-1, 1))
            slices.append((pos, pos + v.size))
            pos += v.size
    return np.concatenate(chunks, axis=0), slices


def vector2params(vector, params):
    """Inverse of params2vector (utils.py:127-136)."""
    out, pos = copy.deepcopy(params), 0
    for li, layer_params in enumerate(params):
        for k, v in layer_params.items():
            v = np.asarray(v)
            out[li][k] = np.reshape(vector[pos:pos + v.size], v.shape)
            pos += v.size
    return out


def sample_params_slices(params_slides, max_samples, flat=True):
    """utils.py:139-148."""
    picked = []
    for lo, hi in params_slides:
        chosen = random.sample(range(lo, hi), min(hi - lo, max_samples))
        if flat:
            picked.extend(chosen)
        else:
            picked.append(chosen)
    return picked
