"""LeakyRelu, Relu, PRelu, Sigmoid, Tanh, Softmax, LogSoftmax (dnpy/layers/activations.py)."""
import numpy as np

from .base import Layer
from .. import initializers
from ..darray import LazyDArray, ParamDict


class LeakyRelu(Layer):
    """activations.py:7-22.  The gate ``x > 0`` is recomputed from the (still resident) input in
    backward instead of being stored: 3 streams of HBM traffic instead of 4."""

    def __init__(self, l_in, alpha=0.3, name="LeakyRelu"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.alpha = alpha
        self.oshape = self.parents[0].oshape
        self.gate = None

    def forward(self):
        if self.__dict__.pop("_cpa_skip_fwd", False):      # the fused (conv->)pool->act kernel wrote self.output already
            return
        x = self._in()
        out = self._buf('out', x.shape)
        self._call("dnb_leakyrelu_fwd", x.ptr(), out.ptr(), x.size, float(self.alpha))
        self.output = out

    def _backward_now(self):
        x, dy = self._in(), self._delta()
        dx = self._buf('dx', x.shape)
        self._call("dnb_leakyrelu_bwd", x.ptr(), dy.ptr(), dx.ptr(), x.size, float(self.alpha))
        return dx

    def backward(self):
        block = getattr(self, "_cpa_block", None)
        if block is not None and block._cpa_live is not None:
            # fused block (layers/convolution.py, layers/pooling.py): the conv / pool layer consumes THIS layer's incoming
            # delta directly; the pool's delta exists only if somebody reads it
            self._delta()
            self.parents[0].delta = LazyDArray(self.parents[0].output.shape, lambda: self._backward_now().dev())
            return
        self.parents[0].delta = self._backward_now()


class Relu(LeakyRelu):
    """activations.py:25-28."""

    def __init__(self, l_in, name="Relu"):
        super().__init__(l_in, alpha=0.0, name=name)


class PRelu(Layer):
    """activations.py:31-66; alpha has the full oshape, galpha += mean_b(delta * (x <= 0))."""

    def __init__(self, l_in, alpha_initializer=None, alpha_regularizer=None, name="LeakyRelu"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.oshape = self.parents[0].oshape
        self.gate = None
        self.params = ParamDict(alpha=np.zeros(self.oshape))
        self.grads = ParamDict(alpha=np.zeros(self.oshape))
        self.alpha_initializer = alpha_initializer if alpha_initializer is not None else initializers.Zeros()
        self.alpha_regularizer = alpha_regularizer

    def initialize(self, optimizer=None):
        super().initialize(optimizer=optimizer)
        self.alpha_initializer.apply(self.params, ['alpha'])

    def forward(self):
        x = self._in()
        out = self._buf('out', x.shape)
        self._call("dnb_prelu_fwd", x.ptr(), self.params['alpha'].ptr(), out.ptr(), x.shape[0],
                   x.size // max(1, x.shape[0]))
        self.output = out

    def backward(self):
        x, dy = self._in(), self._delta()
        dx = self._buf('dx', x.shape)
        b = x.shape[0]
        self._call("dnb_prelu_bwd", x.ptr(), self.params['alpha'].ptr(), dy.ptr(), dx.ptr(), self.grads['alpha'].ptr(),
                   b, x.size // max(1, b), self._inv_m(b))
        self.grads['alpha'].written()
        self.parents[0].delta = dx


class _OutAct(Layer):
    """Activations whose backward needs only their own output."""
    _fwd = _bwd = None

    def __init__(self, l_in, name):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.oshape = self.parents[0].oshape

    def forward(self):
        x = self._in()
        out = self._buf('out', x.shape)
        self._call(self._fwd, x.ptr(), out.ptr(), x.size)
        self.output = out

    def backward(self):
        y, dy = self.output, self._delta()
        dx = self._buf('dx', y.shape)
        self._call(self._bwd, y.ptr(), dy.ptr(), dx.ptr(), y.size)
        self.parents[0].delta = dx


class Sigmoid(_OutAct):
    """activations.py:69-81: 1/(1+exp(-x+eps)+eps) (Q18)."""
    _fwd, _bwd = "dnb_sigmoid_fwd", "dnb_sigmoid_bwd"

    def __init__(self, l_in, name="Sigmoid"):
        super().__init__(l_in, name)


class Tanh(_OutAct):
    """activations.py:84-98."""
    _fwd, _bwd = "dnb_tanh_fwd", "dnb_tanh_bwd"

    def __init__(self, l_in, name="Tanh"):
        super().__init__(l_in, name)


class Softmax(Layer):
    """activations.py:101-135."""

    def __init__(self, l_in, stable=True, name="Softmax"):
        super().__init__(name=name)
        self.parents.append(l_in)
        if len(l_in.oshape) != 1:
            raise ValueError(f"Expected a 1D layer ({self.name})")
        self.oshape = self.parents[0].oshape
        self.stable = stable
        self.ce_loss = False

    def forward(self):
        x = self._in()
        out = self._buf('out', x.shape)
        self._call("dnb_softmax_fwd", x.ptr(), out.ptr(), x.shape[0], x.shape[1], int(bool(self.stable)))
        self.output = out

    def backward(self):
        dy = self._delta()
        if self.ce_loss:            # CrossEntropy(Softmax(z)) shortcut: delta passes through (zero copy)
            self.parents[0].delta = dy
            return
        y = self.output
        dx = self._buf('dx', y.shape)
        self._call("dnb_softmax_bwd", y.ptr(), dy.ptr(), dx.ptr(), y.shape[0], y.shape[1])
        self.parents[0].delta = dx


class LogSoftmax(Layer):
    """activations.py:138-163 (backward passes the delta through)."""

    def __init__(self, l_in, stable=True, name="LogSoftmax"):
        super().__init__(name=name)
        self.parents.append(l_in)
        if len(l_in.oshape) != 1:
            raise ValueError(f"Expected a 1D layer ({self.name})")
        self.oshape = self.parents[0].oshape
        self.stable = stable

    def forward(self):
        x = self._in()
        out = self._buf('out', x.shape)
        self._call("dnb_logsoftmax_fwd", x.ptr(), out.ptr(), x.shape[0], x.shape[1], int(bool(self.stable)))
        self.output = out

    def backward(self):
        self.parents[0].delta = self._delta()
