"""Device-resident arrays that look like ndarrays at the dnpy boundary.

dnpy layers talk through attributes (``layer.output``, ``layer.delta``,
``layer.params[k]``, ``layer.grads[k]`` — dnpy/layers/base.py:12-20) that user
code and the reference's own tests read and assign as NumPy arrays.  ``DArray``
keeps the data in HBM as fp32 (int32 for token indices) and converts only when
somebody on the host looks: ``np.asarray(d)``, ``d == ref``, ``d[...]``,
``d.fill(0.0)`` all work, assignment of an ndarray uploads.

torch is used here for what it is good at — device memory and streams.  No
arithmetic is done with torch: every kernel is in libdnpy_b200.so.
"""
import numpy as np

from . import _lib

_torch = None


def torch():
    global _torch
    if _torch is None:
        import torch as _t
        _torch = _t
    return _torch


def require_cuda():
    t = torch()
    if not t.cuda.is_available():
        raise RuntimeError("dnpy_b200: no CUDA device — the backend has no CPU fallback")
    return t


def stream_ptr():
    return torch().cuda.current_stream().cuda_stream


def dev_empty(shape, int_dtype=False):
    t = require_cuda()
    return t.empty(tuple(int(s) for s in shape), dtype=t.int32 if int_dtype else t.float32, device="cuda")


def dev_zeros(shape, int_dtype=False):
    t = require_cuda()
    return t.zeros(tuple(int(s) for s in shape), dtype=t.int32 if int_dtype else t.float32, device="cuda")


class DArray:
    """Coherent (host float64 | device fp32) pair with lazy transfers."""

    __array_priority__ = 100.0

    def __init__(self, host=None, dev=None, is_int=False):
        self.is_int = bool(is_int)
        self._host = None
        self._dev = None
        self._host_ok = False
        self._dev_ok = False
        self._pinned = False          # device storage is a fixed view (arena slot): never re-allocate
        if host is not None:
            self._set_host(host)
        elif dev is not None:
            self._dev = dev
            self._dev_ok = True
            self.shape = tuple(dev.shape)
        else:
            raise ValueError("DArray needs data")

    # -- construction helpers -------------------------------------------------
    @classmethod
    def wrap(cls, obj, is_int=False):
        """ndarray / list / DArray -> DArray (no copy for a DArray)."""
        if isinstance(obj, DArray):
            return obj
        return cls(host=obj, is_int=is_int)

    @classmethod
    def device(cls, tensor, is_int=False):
        return cls(dev=tensor, is_int=is_int)

    def _set_host(self, host):
        a = np.asarray(host)
        self._host = a.astype(np.int64 if self.is_int else np.float64, copy=True)
        self.shape = tuple(self._host.shape)
        self._host_ok = True
        self._dev_ok = False

    # -- device side ----------------------------------------------------------
    def dev(self):
        """The torch CUDA tensor (uploads a pending host copy)."""
        if not self._dev_ok:
            t = require_cuda()
            src = t.from_numpy(np.ascontiguousarray(self._host.astype(np.int32 if self.is_int else np.float32)))
            if self._dev is None or tuple(self._dev.shape) != self.shape:
                if self._pinned:
                    raise ValueError(f"cannot rebind an arena slot of shape {tuple(self._dev.shape)} to {self.shape}")
                self._dev = src.to("cuda", non_blocking=False)
            else:
                self._dev.copy_(src)
            self._dev_ok = True
        return self._dev

    def ptr(self):
        return self.dev().data_ptr()

    def written(self):
        """A kernel wrote the device copy: the host copy is stale."""
        self._dev_ok = True
        self._host_ok = False
        return self

    def bind(self, view):
        """Move the storage into ``view`` (a slice of a flat arena) keeping the values."""
        if self._host_ok and not self._dev_ok:
            t = torch()
            view.copy_(t.from_numpy(np.ascontiguousarray(self._host.astype(np.float32))).reshape(view.shape))
        else:
            view.copy_(self.dev().reshape(view.shape))
        self._dev = view
        self._dev_ok = True
        self._pinned = True
        return self

    def assign(self, values):
        """In-place overwrite from host data (keeps an arena binding)."""
        a = np.asarray(values)
        if tuple(a.shape) != self.shape and self._pinned:
            raise ValueError(f"shape mismatch: slot {self.shape} vs {tuple(a.shape)}")
        self._set_host(a)
        if self._pinned or self._dev is not None:
            self.dev()          # same device storage, new values (dev() copies into the existing tensor)
        return self

    # -- host side ------------------------------------------------------------
    def host(self):
        if not self._host_ok:
            h = self._dev.detach().cpu().numpy()
            self._host = h.astype(np.int64 if self.is_int else np.float64)
            self._host_ok = True
        return self._host

    def __array__(self, dtype=None, copy=None):
        h = self.host()
        if dtype is not None and h.dtype != dtype:
            return h.astype(dtype)
        return np.array(h) if copy else h

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def dtype(self):
        return np.dtype(np.int64 if self.is_int else np.float64)

    @property
    def T(self):
        return self.host().T

    def __len__(self):
        return self.shape[0]

    def fill(self, value):
        """``grads[k].fill(0.0)`` of Net.reset_grads (net.py:299-302)."""
        if self._dev is not None and self._dev_ok and not self.is_int:
            _lib.call("dnb_fill_f32", self._dev.data_ptr(), self.size, float(value), stream_ptr())
            self._host_ok = False
        else:
            self.host().fill(value)
            self._dev_ok = False

    def copy(self):
        return np.array(self.host())

    def astype(self, dt):
        return self.host().astype(dt)

    def reshape(self, *shape):
        return self.host().reshape(*shape)

    def tolist(self):
        return self.host().tolist()

    def __getitem__(self, key):
        return self.host()[key]

    def __setitem__(self, key, value):
        h = self.host()
        h[key] = value
        self._dev_ok = False
        if self._pinned:
            self.dev()

    def __iter__(self):
        return iter(self.host())

    def __repr__(self):
        return f"DArray(shape={self.shape}, host_ok={self._host_ok}, dev_ok={self._dev_ok})"

    def __reduce__(self):       # pickles (Net.save) and deep copies as a plain ndarray
        return (np.array, (self.host(),))

    def __deepcopy__(self, memo):
        return np.array(self.host())


class LazyDArray(DArray):
    """A tensor nobody on the training path consumes (the delta of an Input layer, the pre-pool output of a fused
    conv->pool block): it is NOT computed when the layer runs; ``thunk()`` launches the kernels that produce it the
    first time somebody looks (``np.asarray``, ``==``, ``.ptr()`` ...) and must only read buffers that stay valid
    until the layer runs again (static layer buffers, parameter snapshots).  ``written()`` — the layer ran again —
    drops the materialised copy."""

    def __init__(self, shape, thunk, is_int=False):
        self.is_int = bool(is_int)
        self._host = None
        self._dev = None
        self._host_ok = False
        self._dev_ok = False
        self._pinned = False
        self.shape = tuple(int(s) for s in shape)
        self._thunk = thunk

    @property
    def materialised(self):
        return self._dev_ok or self._host_ok

    def dev(self):
        if not self.materialised:
            self._dev = self._thunk()
            assert tuple(self._dev.shape) == self.shape, (tuple(self._dev.shape), self.shape)
            self._dev_ok = True
        return super().dev()

    def host(self):
        if not self.materialised:
            self.dev()
        return super().host()

    def written(self):
        self._dev, self._host = None, None
        self._dev_ok = self._host_ok = False
        return self


def _binary(name):
    def op(self, other):
        return getattr(self.host(), name)(np.asarray(other))
    return op


for _n in ("__eq__", "__ne__", "__lt__", "__le__", "__gt__", "__ge__", "__add__", "__sub__", "__mul__",
           "__truediv__", "__radd__", "__rsub__", "__rmul__", "__rtruediv__", "__pow__"):
    setattr(DArray, _n, _binary(_n))
DArray.__hash__ = object.__hash__
DArray.__neg__ = lambda self: -self.host()


class ParamDict(dict):
    """``layer.params`` / ``layer.grads``: values are DArrays; assigning an ndarray
    (initializers rebinding ``params[k]``, examples/9_rnn_sentiment.py:102-110 poking
    weights) uploads into the existing slot when the shape matches."""

    def __init__(self, *a, **kw):
        super().__init__()
        for k, v in dict(*a, **kw).items():
            self[k] = v

    def __setitem__(self, key, value):
        cur = dict.get(self, key)
        if isinstance(value, DArray):
            dict.__setitem__(self, key, value)
        elif cur is not None and tuple(np.shape(value)) == cur.shape:
            # in place, pinned (arena slot) or not (BatchNorm moving stats): captured step graphs hold the buffer's
            # address, so Net.set_params / Net.load must never swap the device storage of an existing parameter
            cur.assign(value)
        else:
            if cur is not None and cur._pinned:
                raise ValueError(f"params['{key}']: cannot change the shape of an arena-bound parameter "
                                 f"({cur.shape} -> {tuple(np.shape(value))})")
            if cur is not None and cur._dev is not None:
                from .layers.base import Layer        # a device buffer is being replaced: captured graphs are stale
                Layer.buffer_generation += 1
            dict.__setitem__(self, key, DArray(host=value))

    def host_copy(self):
        return {k: np.array(v.host()) for k, v in self.items()}
