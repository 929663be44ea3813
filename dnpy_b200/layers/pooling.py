"""MaxPool2D, GlobalMaxPool2D, AvgPool2D, GlobalAvgPool2D (dnpy/layers/pooling.py)."""
import numpy as np

from .base import Layer
from .. import utils, _lib
from ..darray import DArray, LazyDArray
from ..runtime import config
from .._lib import Win


class Pool(Layer):
    """pooling.py:7-28, including Q4: the padding is derived from ``_oshape[1:]`` (width only)."""

    def __init__(self, l_in, pool_size, strides, padding, name="Pool"):
        super().__init__(name=name)
        self.parents.append(l_in)
        input_size = l_in.oshape[1:]
        self.pool_size = pool_size
        self.strides = strides
        self.padding = padding
        _oshape = utils.get_output(input_size=input_size, kernel_size=pool_size, strides=strides,
                                   padding=padding, dilation_rate=None)
        _pads = utils.get_padding(padding=padding, input_size=input_size, output_size=_oshape[1:],
                                  kernel_size=pool_size, strides=strides)
        self.pads = tuple(_pads)
        self.oshape = tuple([l_in.oshape[0]] + _oshape.tolist())
        (self.pad_top, self.pad_bottom), (self.pad_left, self.pad_right) = utils.get_side_paddings(self.pads)

    def _win(self, batch):
        c, h, w = self.parents[0].oshape
        _, oh, ow = self.oshape
        if min(self.pad_top, self.pad_bottom, self.pad_left, self.pad_right) < 0:
            raise ValueError(f"{self.name}: negative padding")
        if (oh - 1) * self.strides[0] + self.pool_size[0] > h + self.pad_top + self.pad_bottom or \
                (ow - 1) * self.strides[1] + self.pool_size[1] > w + self.pad_left + self.pad_right:
            # Q4: for non-square inputs the width-derived padding is too small and NumPy would raise/mis-slice
            raise ValueError(f"{self.name}: pooling windows exceed the padded input (non-square 'same' pooling)")
        return Win(batch, c, h, w, c, oh, ow, self.pool_size[0], self.pool_size[1], self.strides[0], self.strides[1],
                   self.pad_top, self.pad_bottom, self.pad_left, self.pad_right)


class MaxPool2D(Pool):
    """pooling.py:31-99.  ``output_idxs`` keeps the reference's meaning: the row-major offset of
    the first maximum inside the zero-padded window."""

    def __init__(self, l_in, pool_size, strides=(2, 2), padding="none", name="MaxPool2D"):
        super().__init__(l_in, pool_size=pool_size, strides=strides, padding=padding, name=name)
        self.output_idxs = None

    # ---- fused MaxPool2D -> LeakyRelu/Relu pair (dnb_maxpool2d_act_fwd / dnb_pool_act_bwd), planned by Net.build -------
    _pa = None            # the activation, when it is this pool's only consumer and the geometry fits the byte format
    _cpa_live = None      # state of the pair between this step's forward and backward (the name the activation looks for)

    def _pa_usable(self, batch):
        # the byte-tensor backward implements the per-sample routing (identical to the literal one for a single sample)
        return (self._pa is not None and not self.debug and
                (config.route == _lib.ROUTE_PER_SAMPLE or batch == 1))

    def _pair_buffers(self, b):
        """(activation output, byte tensor) the fused forward writes."""
        n = b * int(np.prod(self.oshape))
        return self._pa._buf('out', (b, *self.oshape)), self._raw('pa_idx_gate', n)

    def _install_pair(self, x, aout, ig, b):
        """State after a fused forward wrote ``aout`` / ``ig`` (this layer's kernel, or the epilogue of the tensor-core
        convolution in front of it): the tensors nobody wrote stay observable, computed by the plain kernel on a read."""
        act, g = self._pa, self._win(b)
        n = b * int(np.prod(self.oshape))

        def pool_out():
            out = self._buf('out', (b, *self.oshape))
            idx = self._buf('idx', (b, *self.oshape), is_int=True)
            self._call("dnb_maxpool2d_fwd", x.ptr(), out.ptr(), idx.ptr(), g)
            return out.dev()

        def pool_idx():
            idx = self._buf('idx_unpacked', (b, *self.oshape), is_int=True)
            self._call("dnb_idx_gate_unpack", ig, idx.ptr(), n)
            return idx.dev()
        self.output = LazyDArray((b, *self.oshape), pool_out)
        self.output_idxs = LazyDArray((b, *self.oshape), pool_idx, is_int=True)
        act.output = aout
        act._cpa_skip_fwd = True
        self._cpa_live = dict(ig=ig, idx=self.output_idxs, b=b)

    def _forward_pair(self, x):
        b = x.shape[0]
        aout, ig = self._pair_buffers(b)                   # ig: one byte per pooled element, argmax offset | gate << 4
        self._call("dnb_maxpool2d_act_fwd", x.ptr(), aout.ptr(), ig, self._win(b), float(self._pa.alpha))
        self._install_pair(x, aout, ig, b)

    def forward(self):
        if self.__dict__.pop("_cpa_skip_fwd", False):      # the fused conv->pool->act block set output / output_idxs (lazy)
            return
        if self.__dict__.pop("_pair_done", False):         # the tensor-core convolution in front ran this pair in its epilogue
            return
        x = self._in()
        b = x.shape[0]
        self._cpa_live = None
        if self._pa_usable(b):
            return self._forward_pair(x)
        out = self._buf('out', (b, *self.oshape))
        idx = self._buf('idx', (b, *self.oshape), is_int=True)
        self._call("dnb_maxpool2d_fwd", x.ptr(), out.ptr(), idx.ptr(), self._win(b))
        self.output = out
        self.output_idxs = idx

    def _backward_now(self, route):
        x, dy = self._in(), self._delta()
        b = x.shape[0]
        g = self._win(b)
        dx = self._buf('dx', x.shape)
        ws = self._raw('ws', _lib.load().dnb_maxpool2d_bwd_ws_bytes(g, route))
        self._call("dnb_maxpool2d_bwd", dy.ptr(), DArray.wrap(self.output_idxs, is_int=True).ptr(), dx.ptr(), g, route, ws)
        return dx

    def backward(self):
        route = config.route
        live, self._cpa_live = self._cpa_live, None
        if live is not None and self._pa_usable(live['b']):
            # LeakyRelu.backward + MaxPool2D.backward in one pass over the activation's delta and the byte tensor
            act, b = self._pa, live['b']
            if self.output_idxs is not live['idx']:        # the caller assigned its own argmax tensor: route through it
                user = DArray.wrap(self.output_idxs, is_int=True)
                self._call("dnb_idx_gate_set", user.ptr(), live['ig'], user.size)
            dx = self._buf('dx', (b, *self.parents[0].oshape))
            self._call("dnb_pool_act_bwd", act._delta().ptr(), live['ig'], dx.ptr(), self._win(b), float(act.alpha))
            self.parents[0].delta = dx
            return
        block = getattr(self, "_cpa_block", None)
        if block is not None and block._cpa_live is not None:
            # fused block: the conv layer computes its gradients from the activation's delta; this layer's 4.7x larger
            # delta (the conv layer's dY) is only materialised if somebody reads it
            shape = (self.output.shape[0], *self.parents[0].oshape)
            self.parents[0].delta = LazyDArray(shape, lambda: self._backward_now(route).dev())
            return
        self.parents[0].delta = self._backward_now(route)


class GlobalMaxPool2D(MaxPool2D):
    """pooling.py:102-106."""

    def __init__(self, l_in, name="GlobalMaxPool2D"):
        super().__init__(l_in, pool_size=l_in.oshape[1:], strides=(1, 1), padding="none", name=name)


class AvgPool2D(Pool):
    """pooling.py:109-174 (backward only exists for 4-element windows, as in the reference)."""

    def __init__(self, l_in, pool_size, strides=(2, 2), padding="none", name="AvgPool2D"):
        super().__init__(l_in, pool_size=pool_size, strides=strides, padding=padding, name=name)

    def forward(self):
        x = self._in()
        b = x.shape[0]
        out = self._buf('out', (b, *self.oshape))
        self._call("dnb_avgpool2d_fwd", x.ptr(), out.ptr(), self._win(b))
        self.output = out

    def backward(self):
        x, dy = self._in(), self._delta()
        dx = self._buf('dx', x.shape)
        self._call("dnb_avgpool2d_bwd", dy.ptr(), dx.ptr(), self._win(x.shape[0]))
        self.parents[0].delta = dx


class GlobalAvgPool2D(AvgPool2D):
    """pooling.py:177-181."""

    def __init__(self, l_in, name="GlobalAvgPool2D"):
        super().__init__(l_in, pool_size=l_in.oshape[1:], strides=(1, 1), padding="none", name=name)
