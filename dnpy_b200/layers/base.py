"""Layer protocol — the drop-in boundary (dnpy/layers/base.py:6-97).

Same attributes and methods as the reference's ``Layer`` (``parents, output,
delta, oshape, params, grads, training, frozen, optimizer, epsilon``;
``initialize / forward / backward``), but ``output``/``delta``/``params``/
``grads`` hold ``DArray`` handles to HBM.  Each layer keeps its own cached
device buffers (static addresses, so a whole step can be captured in a CUDA
graph) and hands its parent a DArray for ``parent.delta`` — assigned, never
accumulated, exactly like the reference.
"""
import copy

import numpy as np

from .. import _lib
from ..darray import DArray, ParamDict, dev_empty, stream_ptr
from ..runtime import config


class Layer:
    # bumped whenever any layer (re)allocates a device buffer: captured CUDA graphs hold raw addresses and are
    # only replayed while this generation is unchanged
    buffer_generation = 0

    def __init__(self, name):
        self.name = name
        self.training = False

        self.parents = []
        self.output = None
        self.delta = None

        self.oshape = None
        self.batch_size = None

        self.params = ParamDict()
        self.grads = ParamDict()

        self.epsilon = 10e-8
        self.frozen = False
        self.optimizer = None

        self.index = 0
        self.debug = False
        self._bufs = {}
        # set by Net.build for a layer whose data gradient only feeds an Input layer: nothing on the training path
        # consumes it, so backward skips it and hands the Input a LazyDArray (computed when somebody reads it)
        self._lazy_dx = False

    def __str__(self):
        return self.name

    def __setattr__(self, key, value):
        # ``layer.params = {...}`` (Net.set_params, net.py:422) must stay a ParamDict
        if key in ("params", "grads") and not isinstance(value, ParamDict):
            value = ParamDict(value)
        object.__setattr__(self, key, value)

    def initialize(self, optimizer=None):
        """base.py:31-35: every layer owns a deep copy of the optimizer."""
        if optimizer:
            self.optimizer = copy.deepcopy(optimizer)
            self.optimizer.initialize(self.params)

    def forward(self):
        pass

    def backward(self):
        pass

    def get_num_params(self):
        """base.py:43-52."""
        tr, notr = 0, 0
        for k, v in self.params.items():
            if k in self.grads:
                tr += v.size
            else:
                notr += v.size
        return tr, notr

    def print_stats(self, print_tensors=False):
        """base.py:54-73 (debug dump; forces D2H copies)."""
        def line(tag, v):
            a = np.asarray(v)
            print(f"\t\t [{tag}]\tshape={a.shape}; max={float(np.max(a))}; min={float(np.min(a))}; avg={float(np.mean(a))}")
            if print_tensors:
                print(a)
        print(f"\t=> [DEBUG]: {self.name} layer:")
        if self.parents and self.parents[0] is not None and self.parents[0].output is not None:
            line("input", self.parents[0].output)
        if self.output is not None:
            line("output", self.output)
        if self.delta is not None:
            line("delta", self.delta)
        for k in self.params.keys():
            line(k, self.params[k])
        for k in self.grads.keys():
            line(k, self.grads[k])

    # -- device plumbing --------------------------------------------------------------------
    def _in(self, i=0, is_int=False):
        """Parent i's output as a DArray (an ndarray fed by the user is uploaded once)."""
        p = self.parents[i]
        if not isinstance(p.output, DArray):
            if p.output is None:
                raise ValueError(f"{self.name}: parent '{p.name}' has no output yet")
            p.output = DArray(host=p.output, is_int=is_int)
        return p.output

    def _delta(self):
        if not isinstance(self.delta, DArray):
            if self.delta is None:
                raise ValueError(f"{self.name}: no delta to back-propagate")
            self.delta = DArray(host=self.delta)
        return self.delta

    def _buf(self, key, shape, is_int=False):
        """Cached device buffer wrapped in a (stable) DArray."""
        shape = tuple(int(s) for s in shape)
        cur = self._bufs.get(key)
        if cur is None or cur.shape != shape:
            cur = DArray.device(dev_empty(shape, int_dtype=is_int), is_int=is_int)
            self._bufs[key] = cur
            Layer.buffer_generation += 1
        return cur.written()

    def _raw(self, key, nbytes):
        """Cached raw workspace (returns a device pointer or None for 0 bytes)."""
        if nbytes <= 0:
            return None
        n = (int(nbytes) + 3) // 4
        cur = self._bufs.get(key)
        if cur is None or cur.size < n:
            cur = DArray.device(dev_empty((n,)))
            self._bufs[key] = cur
            Layer.buffer_generation += 1
        return cur.ptr()

    def _snapshot(self, key, src):
        """Copy of a small tensor (weights) taken NOW, for a lazy result that may be read after the optimizer ran."""
        snap = self._buf(key, src.shape)
        snap.dev().copy_(src.dev())
        return snap

    @staticmethod
    def _call(name, *args):
        _lib.call(name, *args, stream_ptr())

    @staticmethod
    def _inv_m(batch):
        """1/m of the reference's batch means, with the data-parallel 1/world folded in."""
        return config.inv_world / float(batch)


class MLayer(Layer):
    """Multi-input layer (base.py:76-97)."""

    def __init__(self, l_in, name):
        super().__init__(name=name)
        if not isinstance(l_in, list):
            raise ValueError("A list of layers is expected")
        if len(l_in) < 2:
            raise ValueError("A minimum of two inputs is expected")
        first = l_in[0].oshape
        for other in l_in[1:]:
            if first != other.oshape:
                raise ValueError(f"Layers with different dimensions: {first} vs {other.oshape}")
        self.parents = l_in
        self.oshape = self.parents[0].oshape
