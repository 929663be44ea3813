"""Merge layers (dnpy/layers/merge.py).  ``Add`` is on the hot path (config 4); the other
point-wise merges have no config and no kernel yet."""
import ctypes as C

from .base import MLayer


class MPWLayer(MLayer):
    _kernel = None

    def __init__(self, l_in, operator=None, name="MPWLayer"):
        super().__init__(l_in, name=name)
        self.operator = operator

    def forward(self):
        if self._kernel is None:
            raise NotImplementedError(f"{type(self).__name__}: no CUDA kernel for this merge (out of scope, SURVEY §8f)")
        xs = [self._in(i) for i in range(len(self.parents))]
        out = self._buf('out', xs[0].shape)
        ptrs = (C.c_void_p * len(xs))(*[x.ptr() for x in xs])
        self._call(self._kernel, ptrs, len(xs), out.ptr(), out.size)
        self.output = out

    def backward(self):
        """merge.py:18-20: the SAME delta object goes to every parent (Q10) — zero copy."""
        d = self._delta()
        for p in self.parents:
            p.delta = d


class Add(MPWLayer):
    _kernel = "dnb_add_n"

    def __init__(self, l_in, name="Add"):
        super().__init__(l_in, name=name)


def _unsupported(cls_name):
    class _M(MPWLayer):
        def __init__(self, l_in, name=cls_name):
            super().__init__(l_in, name=name)
    _M.__name__ = cls_name
    return _M


Subtract = _unsupported("Subtract")
Multiply = _unsupported("Multiply")
Divide = _unsupported("Divide")
Power = _unsupported("Power")
Maximum = _unsupported("Maximum")
Minimum = _unsupported("Minimum")
