"""Dropout, GaussianNoise, BatchNorm (dnpy/layers/regularization.py)."""
import numpy as np

from .base import Layer
from .. import initializers
from ..darray import DArray, ParamDict


def _global_draw(draw, shape):
    """One host draw for the GLOBAL minibatch from NumPy's global stream.  Under data parallelism every rank holds the
    same stream state (identical seeds), draws the full (B_local * world, ...) block exactly like a single process
    would (regularization.py:19,41) and keeps its own contiguous row block."""
    from ..runtime import config
    shape = tuple(int(s) for s in shape)
    if config.world == 1:
        return draw(shape)
    full = draw((shape[0] * config.world,) + shape[1:])
    return full[config.rank * shape[0]:(config.rank + 1) * shape[0]]


class _HostDrawLayer(Layer):
    """Layers whose training forward consumes the host RNG (Dropout, GaussianNoise): the draw is made by ``prepare``
    into a STATIC device buffer, so a captured step graph replays with fresh values.  Net calls ``prepare`` before it
    captures / replays a step graph; an eager forward calls it itself."""

    def _draw(self, shape):
        raise NotImplementedError

    def prepare(self, shape):
        shape = tuple(int(s) for s in shape)
        buf = self._bufs.get('draw')
        if buf is None or buf.shape != shape:
            from ..darray import dev_empty
            buf = DArray.device(dev_empty(shape))
            self._bufs['draw'] = buf
            Layer.buffer_generation += 1
        import torch
        host = np.ascontiguousarray(_global_draw(self._draw, shape), dtype=np.float32)
        buf.dev().copy_(torch.from_numpy(host))
        buf.written()
        self._prepared = True
        return buf

    def _drawn(self, shape):
        if not getattr(self, "_prepared", False):
            self.prepare(shape)
        self._prepared = False
        return self._bufs['draw']


class Dropout(_HostDrawLayer):
    """regularization.py:7-25.  Non-inverted dropout (Q12).  The mask is drawn on the host from
    NumPy's GLOBAL stream — one ``np.random.random(shape)`` per training forward, exactly the
    reference's call — so masks are bit-identical for a shared seed; only the 0/1 gate is uploaded."""

    def __init__(self, l_in, rate=0.5, name="Dropout"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.oshape = self.parents[0].oshape
        self.rate = rate
        self.gate = None

    def _draw(self, shape):
        return (np.random.random(shape) > self.rate).astype(np.float32)

    def forward(self):
        x = self._in()
        out = self._buf('out', x.shape)
        if self.training:
            self.gate = self._drawn(x.shape)
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


class GaussianNoise(_HostDrawLayer):
    """regularization.py:28-47: host-drawn noise (same global stream, one ``np.random.normal`` per training forward),
    added on the device from a static buffer (graph-capturable, like Dropout's gate)."""

    def __init__(self, l_in, mean=0.0, stddev=1.0, name="GaussianNoise"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.mean = mean
        self.stddev = stddev
        self.oshape = self.parents[0].oshape

    def _draw(self, shape):
        return np.random.normal(loc=self.mean, scale=self.stddev, size=shape)

    def forward(self):
        x = self._in()
        if not self.training:
            self.output = x
            return
        import ctypes as C
        noise = self._drawn(x.shape)
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
        from ..runtime import config
        if self.training and config.world > 1:
            # the statistics of the GLOBAL minibatch (regularization.py:98-99 over all ranks' rows): every rank's
            # (mean, M2) lands in its row of `slots`, a SUM leaves all rows everywhere, the kernel merges them
            from .. import parallel
            slots = self._buf('stat_slots', (config.world, 2, n_feat))
            slots.dev().zero_()
            mine = slots.dev()[config.rank]
            self._call("dnb_batchnorm_stats", x.ptr(), mine[0].data_ptr(), mine[1].data_ptr(), b, n_feat)
            parallel.gather_slots(slots.dev())
            self._call("dnb_batchnorm_fwd_synced", x.ptr(), self.params['gamma'].ptr(), self.params['beta'].ptr(),
                       mm.ptr(), mv.ptr(), slots.ptr(), config.world, mu.ptr(), std.ptr(), out.ptr(), b, n_feat,
                       float(self.momentum), float(self.epsilon))
        else:
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
        from ..runtime import config
        if config.world > 1:
            from .. import parallel
            sums = self._buf('bwd_sums', (2, n_feat))
            self._call("dnb_batchnorm_bwd_sums", x.ptr(), self.cache['mu'].ptr(), self.cache['inv_var'].ptr(), dy.ptr(),
                       sums.ptr(), self.grads['gamma'].ptr(), self.grads['beta'].ptr(), b, n_feat)
            parallel.allreduce(sums.dev(), "sum")
            self._call("dnb_batchnorm_bwd_synced", x.ptr(), self.params['gamma'].ptr(), self.cache['mu'].ptr(),
                       self.cache['inv_var'].ptr(), dy.ptr(), sums.ptr(), dx.ptr(), b, b * config.world, n_feat)
        else:
            self._call("dnb_batchnorm_bwd", x.ptr(), self.params['gamma'].ptr(), self.cache['mu'].ptr(),
                       self.cache['inv_var'].ptr(), dy.ptr(), dx.ptr(), self.grads['gamma'].ptr(),
                       self.grads['beta'].ptr(), b, n_feat)
        self.grads['gamma'].written()
        self.grads['beta'].written()
        self.parents[0].delta = dx
