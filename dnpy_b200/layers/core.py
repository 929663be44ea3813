"""Input, Embedding, Dense, Reshape, Flatten (dnpy/layers/core.py)."""
import numpy as np

from .base import Layer
from .. import initializers
from ..darray import DArray, LazyDArray, ParamDict
from ..regularizers import reg_coeffs
from ..runtime import config
from .._lib import Reg


class Input(Layer):
    """core.py:7-18 — the H2D boundary: ``l_in.output = ndarray`` is uploaded by the first consumer."""

    def __init__(self, shape, batch_size=None, name="Input"):
        super().__init__(name=name)
        self.oshape = shape
        self.batch_size = batch_size


class Embedding(Layer):
    """core.py:21-65. input_dim: vocabulary, output_dim: latent size, input_length: tokens per sample."""

    def __init__(self, l_in, input_dim, output_dim, input_length, embeddings_initializer=None,
                 embeddings_regularizer=None, name="Embedding"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.oshape = (input_length, output_dim)
        self.params = ParamDict(w1=np.zeros((input_dim, output_dim)))
        self.grads = ParamDict(w1=np.zeros((input_dim, output_dim)))
        self.embeddings_initializer = embeddings_initializer if embeddings_initializer is not None \
            else initializers.RandomUniform()
        self.embeddings_regularizer = embeddings_regularizer

    def initialize(self, optimizer=None):
        super().initialize(optimizer=optimizer)
        self.embeddings_initializer.apply(self.params, ['w1'])

    def forward(self):
        idx = self._in(0, is_int=True)
        assert idx.ndim == 2
        b, l = idx.shape
        v, d = self.params['w1'].shape
        out = self._buf('out', (b, l, d))
        self._call("dnb_embedding_fwd", idx.ptr(), self.params['w1'].ptr(), out.ptr(), b, l, v, d)
        self.output = out

    def backward(self):
        """core.py:62-65: batch mean, then last-write-wins scatter (Q15); no delta for the parent."""
        idx = self._in(0, is_int=True)
        dy = self._delta()
        b, l = idx.shape
        v, d = self.params['w1'].shape
        from .. import _lib
        nbytes = _lib.load().dnb_embedding_bwd_ws_bytes(l, v, d)
        if config.world > 1:
            # exact under data parallelism: the batch mean over ALL ranks' rows (SUM of the local parts) and the
            # last GLOBAL position of every vocabulary row (MAX); every rank then adds 1/world of the exact gradient, which
            # the gradient all-reduce sums back
            from .. import parallel
            import torch
            wsb = self._buf('ws_dp', ((nbytes + 3) // 4,))
            raw = wsb.dev()
            self._call("dnb_embedding_bwd_prepare", idx.ptr(), dy.ptr(), b, l, v, d, self._inv_m(b),
                       config.rank * b * l, wsb.ptr())
            parallel.allreduce(raw[:v].view(torch.int32), "max")
            parallel.allreduce(raw[v:v + l * d], "sum")
            self._call("dnb_embedding_bwd_apply", wsb.ptr(), self.grads['w1'].ptr(), l, v, d, config.inv_world)
        else:
            ws = self._raw('ws', nbytes)
            self._call("dnb_embedding_bwd", idx.ptr(), dy.ptr(), self.grads['w1'].ptr(), b, l, v, d,
                       self._inv_m(b), ws)
        self.grads['w1'].written()


class Dense(Layer):
    """core.py:68-135."""

    def __init__(self, l_in, units, kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, bias_regularizer=None, name="Dense"):
        super().__init__(name=name)
        self.parents.append(l_in)
        if len(l_in.oshape) != 1:
            raise ValueError(f"Expected a 1D layer ({self.name})")
        if units == -1:
            units = l_in.oshape[0]
        self.oshape = (units,)
        self.units = units
        fan_in = self.parents[0].oshape[0]
        self.params = ParamDict(w1=np.zeros((fan_in, units)), b1=np.zeros((1, units)))
        self.grads = ParamDict(w1=np.zeros((fan_in, units)), b1=np.zeros((1, units)))
        self.kernel_initializer = kernel_initializer if kernel_initializer is not None \
            else initializers.HeNormal(fan_in=fan_in, fan_out=units)
        # Q3 (core.py:100-103): a given bias_initializer is ignored, the KERNEL initializer is used
        self.bias_initializer = initializers.Zeros() if bias_initializer is None else kernel_initializer
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer

    def initialize(self, optimizer=None):
        super().initialize(optimizer=optimizer)
        self.kernel_initializer.apply(self.params, ['w1'])
        self.bias_initializer.apply(self.params, ['b1'])

    def forward(self):
        x = self._in()
        b, fin = x.shape
        out = self._buf('out', (b, self.units))
        self._call("dnb_dense_fwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), out.ptr(),
                   b, fin, self.units, config.math)
        self.output = out

    def backward(self):
        """core.py:119-135: dX assigned to the parent (even an Input), grads accumulated with /m."""
        x = self._in()
        dy = self._delta()
        b, fin = x.shape
        lazy = self._lazy_dx and not self.debug
        dx = None if lazy else self._buf('dx', (b, fin))
        self._call("dnb_dense_bwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), dy.ptr(),
                   None if lazy else dx.ptr(), self.grads['w1'].ptr(), self.grads['b1'].ptr(), b, fin, self.units,
                   self._inv_m(b), Reg(*reg_coeffs(self.kernel_regularizer)), Reg(*reg_coeffs(self.bias_regularizer)),
                   config.reg_scale, config.math)
        self.grads['w1'].written()
        self.grads['b1'].written()
        if lazy:
            # only an Input layer would receive this delta: dead on the training path, computed on demand from this
            # step's dy buffer and a snapshot of the weights
            w_now, math, units = self._snapshot('w_snap', self.params['w1']), config.math, self.units

            def materialise():
                out = self._buf('dx', (b, fin))
                self._call("dnb_dense_bwd_data", w_now.ptr(), dy.ptr(), out.ptr(), b, fin, units, math)
                return out.dev()
            dx = LazyDArray((b, fin), materialise)
        self.parents[0].delta = dx


class Reshape(Layer):
    """core.py:138-166 — zero-copy views of the parent's buffers."""

    def __init__(self, l_in, shape, include_batch=False, name="Reshape"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.include_batch = include_batch
        if include_batch:
            self.oshape = shape[1:]
            self.oshape_batch = shape
        else:
            if shape == -1 or shape[0] == -1:
                shape = (int(np.prod(self.parents[0].oshape)),)
            if np.prod(self.parents[0].oshape) != np.prod(shape):
                raise ValueError(f"Not compatible shapes ({self.name})")
            self.oshape = shape
            self.oshape_batch = None

    def forward(self):
        x = self._in()
        new_shape = (-1, *self.oshape) if not self.include_batch else self.oshape_batch
        self.output = DArray.device(x.dev().view(*[int(s) for s in new_shape]))

    def backward(self):
        dy = self._delta()
        new_shape = (-1, *self.parents[0].oshape) if not self.include_batch else self.oshape_batch
        new_shape = [int(s) for s in new_shape]
        if isinstance(dy, LazyDArray) and not dy.materialised:        # stay lazy: a view of a delta nobody computed yet
            full = [int(s) for s in new_shape]
            if full[0] == -1:
                full[0] = dy.size // max(1, int(np.prod(full[1:])))
            self.parents[0].delta = LazyDArray(full, lambda: dy.dev().view(*new_shape))
            return
        self.parents[0].delta = DArray.device(dy.dev().view(*new_shape))


class Flatten(Reshape):
    """core.py:169-172."""

    def __init__(self, l_in, name="Flatten"):
        super().__init__(l_in, shape=(-1,), name=name)
