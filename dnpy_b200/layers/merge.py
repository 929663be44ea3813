"""Merge layers (dnpy/layers/merge.py).  ``Add`` is on the hot path (config 4); the other point-wise merges and
``Average`` share one n-ary kernel (`dnb_merge_n`, a left fold in the reference's order).
``Concatenate`` ("Not tested" in the reference) has no kernel."""
import ctypes as C

from .base import MLayer

_OPS = dict(add=0, subtract=1, multiply=2, divide=3, power=4, maximum=5, minimum=6)


class MPWLayer(MLayer):
    _kernel = None
    _op = None

    def __init__(self, l_in, operator=None, name="MPWLayer"):
        super().__init__(l_in, name=name)
        self.operator = operator

    def forward(self):
        if self._kernel is None and self._op is None:
            raise NotImplementedError(f"{type(self).__name__}: no CUDA kernel for this merge (out of scope, SURVEY §8f)")
        xs = [self._in(i) for i in range(len(self.parents))]
        out = self._buf('out', xs[0].shape)
        ptrs = (C.c_void_p * len(xs))(*[x.ptr() for x in xs])
        if self._kernel is not None:
            self._call(self._kernel, ptrs, len(xs), out.ptr(), out.size)
        else:
            self._call("dnb_merge_n", ptrs, len(xs), out.ptr(), out.size, _OPS[self._op], 1.0)
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


def _pointwise(cls_name, op):
    class _M(MPWLayer):
        _op = op

        def __init__(self, l_in, name=cls_name):
            super().__init__(l_in, name=name)
    _M.__name__ = _M.__qualname__ = cls_name
    _M.__doc__ = f"merge.py: MPWLayer(np.{op}) — left fold over the parents; backward hands the delta to every parent."
    return _M


Subtract = _pointwise("Subtract", "subtract")
Multiply = _pointwise("Multiply", "multiply")
Divide = _pointwise("Divide", "divide")
Power = _pointwise("Power", "power")
Maximum = _pointwise("Maximum", "maximum")
Minimum = _pointwise("Minimum", "minimum")


class Average(MLayer):
    """merge.py:65-79: out = sum(parents) / n; every parent receives delta / n (a fresh array per parent)."""

    def __init__(self, l_in, name="Average"):
        super().__init__(l_in, name=name)

    def forward(self):
        xs = [self._in(i) for i in range(len(self.parents))]
        out = self._buf('out', xs[0].shape)
        ptrs = (C.c_void_p * len(xs))(*[x.ptr() for x in xs])
        self._call("dnb_merge_n", ptrs, len(xs), out.ptr(), out.size, _OPS["add"], 1.0 / float(len(xs)))
        self.output = out

    def backward(self):
        d = self._delta()
        n = len(self.parents)
        for i, p in enumerate(self.parents):
            dx = self._buf(f'dx{i}', d.shape)
            self._call("dnb_scale", d.ptr(), dx.ptr(), d.size, 1.0 / float(n))
            p.delta = dx


class Concatenate(MLayer):
    """merge.py:81-100, marked "Not tested" in the reference — and it is not usable there: the layer keeps the FIRST
    parent's ``oshape`` although its output is wider, and ``backward`` raises ``TypeError`` (``np.take`` with a slice;
    probed against the reference in this container).  Kept for API presence; scoped out with the reference's own failure."""

    def __init__(self, l_in, axis=0, name="Concatenate"):
        super().__init__(l_in, name=name)
        self.axis = axis + 1

    def forward(self):
        raise NotImplementedError("Concatenate: the reference's own backward raises TypeError (merge.py:95-100); out of scope")

    def backward(self):
        raise TypeError("int() argument must be a string, a bytes-like object or a real number, not 'slice'")
