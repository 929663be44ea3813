"""Dropout, GaussianNoise, BatchNorm (dnpy/layers/regularization.py)."""
import numpy as np

from .base import Layer
from .. import initializers
from ..darray import DArray, ParamDict


class Dropout(Layer):
    """regularization.py:7-25.  Non-inverted dropout (Q12).  The mask is drawn on the host from
    NumPy's GLOBAL stream — one ``np.random.random(shape)`` per training forward, exactly the
    reference's call — so masks are bit-identical for a shared seed; only the 0/1 gate is uploaded."""

    def __init__(self, l_in, rate=0.5, name="Dropout"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.oshape = self.parents[0].oshape
        self.rate = rate
        self.gate = None

    def prepare(self, shape):
        """Draw this step's mask from the global NumPy stream (the reference's one call per training forward,
        regularization.py:19) and upload it into the layer's STATIC gate buffer.  Net calls this before it
        replays a captured step graph; an eager forward calls it itself."""
        gate = self._bufs.get('gate')
        if gate is None or gate.shape != tuple(shape):
            from ..darray import dev_empty
            gate = DArray.device(dev_empty(tuple(shape)))
            self._bufs['gate'] = gate
            Layer.buffer_generation += 1
        import torch
        host = (np.random.random(shape) > self.rate).astype(np.float32)
        gate.dev().copy_(torch.from_numpy(host))
        gate.written()
        self.gate = gate
        self._prepared = True

    def forward(self):
        x = self._in()
        out = self._buf('out', x.shape)
        if self.training:
            if not getattr(self, "_prepared", False):
                self.prepare(x.shape)
            self._prepared = False
            self._call("dnb_mul", x.ptr(), self.gate.ptr(), out.ptr(), x.size)
        else:
            self._call("dnb_scale", x.ptr(), out.ptr(), x.size, float(1 - self.rate))
        self.output = out

    def backward(self):
        dy = self._delta()
        if self.gate is None:
            raise TypeError("Dropout.backward before a training forward (gate is None)")
        dx = self._buf('dx', dy.shape)
        self._call("dnb_mul", dy.ptr(), DArray.wrap(self.gate).ptr(), dx.ptr(), dy.size)
        self.parents[0].delta = dx


class GaussianNoise(Layer):
    """regularization.py:28-47: host-drawn noise (same global stream), added on the device."""

    def __init__(self, l_in, mean=0.0, stddev=1.0, name="GaussianNoise"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.mean = mean
        self.stddev = stddev
        self.oshape = self.parents[0].oshape

    def forward(self):
        x = self._in()
        if not self.training:
            self.output = x
            return
        import ctypes as C
        noise = DArray(host=np.random.normal(loc=self.mean, scale=self.stddev, size=x.shape))
        out = self._buf('out', x.shape)
        ptrs = (C.c_void_p * 2)(x.ptr(), noise.ptr())
        self._call("dnb_add_n", ptrs, 2, out.ptr(), x.size)
        self.output = out

    def backward(self):
        self.parents[0].delta = self._delta()


class BatchNorm(Layer):
    """regularization.py:50-156, literal: statistics over axis 0 only (Q8), (C,H,W)-shaped
    gamma/beta/moving stats, and the reference's backward formulas (Q7)."""

    def __init__(self, l_in, momentum=0.99, bias_correction=False, gamma_initializer=None,
                 beta_initializer=None, name="BatchNorm"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.oshape = self.parents[0].oshape
        shp = tuple(self.parents[0].oshape)
        self.params = ParamDict(gamma=np.ones(shp), beta=np.zeros(shp), moving_mu=np.zeros(shp), moving_var=np.ones(shp))
        self.grads = ParamDict(gamma=np.zeros(shp), beta=np.zeros(shp))
        self.cache = {}
        self.fw_steps = 0
        self.momentum = momentum
        self.bias_correction = bias_correction
        if bias_correction:
            raise NotImplementedError("BatchNorm(bias_correction=True) is out of scope (marked 'Not working' in the reference)")
        self.gamma_initializer = gamma_initializer if gamma_initializer is not None else initializers.Ones()
        self.beta_initializer = beta_initializer if beta_initializer is not None else initializers.Zeros()

    def initialize(self, optimizer=None):
        super().initialize(optimizer=optimizer)
        self.gamma_initializer.apply(self.params, ['gamma'])
        self.beta_initializer.apply(self.params, ['beta'])

    def _stat(self, key, n_feat):
        """moving stats may come back from a reference pickle as (1, *oshape): same data, flat view."""
        p = self.params[key]
        if p.size != n_feat:
            raise ValueError(f"{self.name}: params['{key}'] has {p.size} elements, expected {n_feat}")
        return p

    def forward(self):
        x = self._in()
        b = x.shape[0]
        n_feat = x.size // max(1, b)
        out = self._buf('out', x.shape)
        mu = self._buf('mu', (n_feat,))
        std = self._buf('std', (n_feat,))
        if self.training:
            self.fw_steps += 1
        mm, mv = self._stat('moving_mu', n_feat), self._stat('moving_var', n_feat)
        self._call("dnb_batchnorm_fwd", x.ptr(), self.params['gamma'].ptr(), self.params['beta'].ptr(),
                   mm.ptr(), mv.ptr(), mu.ptr(), std.ptr(), out.ptr(), b, n_feat,
                   float(self.momentum), float(self.epsilon), int(bool(self.training)))
        if self.training:
            mm.written()
            mv.written()
        self.cache = {'mu': mu, 'inv_var': std}
        self.output = out

    def backward(self):
        x, dy = self._in(), self._delta()
        b = x.shape[0]
        n_feat = x.size // max(1, b)
        dx = self._buf('dx', x.shape)
        self._call("dnb_batchnorm_bwd", x.ptr(), self.params['gamma'].ptr(), self.cache['mu'].ptr(),
                   self.cache['inv_var'].ptr(), dy.ptr(), dx.ptr(), self.grads['gamma'].ptr(),
                   self.grads['beta'].ptr(), b, n_feat)
        self.grads['gamma'].written()
        self.grads['beta'].written()
        self.parents[0].delta = dx
