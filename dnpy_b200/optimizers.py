"""SGD (dampened-EMA momentum), RMSProp, Adam (dnpy/optimizers.py) as fused single-pass kernels.

Same protocol as the reference: ``initialize(params)`` allocates zero state for EVERY params key
(Q14), ``apply(params, grads, step_i=None)`` skips keys without a gradient.  One deep copy lives
in every layer (base.py:34).  When a Net has bound its parameters to the flat arenas the state
tensors are arena slices too and ``Net.apply_grads`` issues ONE launch for all layers; the
per-key path below is what a standalone layer / user code gets.
"""
import numpy as np

from . import _lib
from .darray import DArray, stream_ptr


def _as_darray(d, k):
    v = d[k]
    if not isinstance(v, DArray):
        v = DArray(host=v)
        d[k] = v
    return v


class Optimizer:
    def __init__(self, name="Base optimizer"):
        self.name = name

    def initialize(self, params):
        pass

    def apply(self, params, grads, step_i=None):
        pass

    @staticmethod
    def _zeros_like(params):
        return {k: DArray(host=np.zeros(np.shape(v))) for k, v in params.items()}


class SGD(Optimizer):
    """optimizers.py:15-55."""

    def __init__(self, lr=0.001, momentum=0.0, nesterov=False, bias_correction=False):
        super().__init__(name="Momentum")
        self.lr = lr
        self.momentum = momentum
        self.nesterov = nesterov
        self.bias_correction = bias_correction
        self.V = {}

    def initialize(self, params):
        self.V = self._zeros_like(params)

    def launch(self, p, g, v, n):
        _lib.call("dnb_opt_sgd", p, g, v, n, float(self.lr), float(self.momentum), int(bool(self.nesterov)), stream_ptr())

    def apply(self, params, grads, step_i=None):
        if self.bias_correction and self.momentum != 0.0:
            # optimizers.py:44 computes lr**step_i with step_i=None from Net.apply_grads -> TypeError
            raise TypeError("unsupported operand type(s) for ** or pow(): 'float' and 'NoneType'")
        for k in list(params.keys()):
            if k not in grads:
                continue
            p, g, v = _as_darray(params, k), _as_darray(grads, k), self.V[k]
            self.launch(p.ptr(), g.ptr(), v.ptr(), p.size)
            p.written(); v.written()


class RMSProp(Optimizer):
    """optimizers.py:58-84."""

    def __init__(self, lr=0.001, rho=0.9, epsilon=10e-8):
        super().__init__(name="RMSProp")
        self.lr = lr
        self.rho = rho
        self.epsilon = epsilon
        self.S = {}

    def initialize(self, params):
        self.S = self._zeros_like(params)

    def launch(self, p, g, s, n):
        _lib.call("dnb_opt_rmsprop", p, g, s, n, float(self.lr), float(self.rho), float(self.epsilon), stream_ptr())

    def apply(self, params, grads, step_i=None):
        for k in list(params.keys()):
            if k not in grads:
                continue
            p, g, s = _as_darray(params, k), _as_darray(grads, k), self.S[k]
            self.launch(p.ptr(), g.ptr(), s.ptr(), p.size)
            p.written(); s.written()


class Adam(Optimizer):
    """optimizers.py:87-126: no bias correction, epsilon under the sqrt (Q13)."""

    def __init__(self, lr=0.001, beta_1=0.9, beta_2=0.999, epsilon=10e-8, bias_correction=False):
        super().__init__(name="Adam")
        self.lr = lr
        self.beta_1 = beta_1
        self.beta_2 = beta_2
        self.epsilon = epsilon
        self.bias_correction = bias_correction
        self.V = {}
        self.S = {}

    def initialize(self, params):
        self.V = self._zeros_like(params)
        self.S = self._zeros_like(params)

    def launch(self, p, g, v, s, n):
        _lib.call("dnb_opt_adam", p, g, v, s, n, float(self.lr), float(self.beta_1), float(self.beta_2),
                  float(self.epsilon), stream_ptr())

    def apply(self, params, grads, step_i=None):
        if self.bias_correction:
            raise TypeError("unsupported operand type(s) for ** or pow(): 'float' and 'NoneType'")
        for k in list(params.keys()):
            if k not in grads:
                continue
            p, g = _as_darray(params, k), _as_darray(grads, k)
            v, s = self.V[k], self.S[k]
            self.launch(p.ptr(), g.ptr(), v.ptr(), s.ptr(), p.size)
            p.written(); v.written(); s.written()
