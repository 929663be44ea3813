"""Conv2D, PointwiseConv2D, DepthwiseConv2D (dnpy/layers/convolution.py).

Shape logic (output size first, then "same" padding, floor-half before / rest
after) is the reference's, computed once on the host.  The arithmetic is the
CUDA implicit convolution behind dnb_conv2d_* / dnb_dwconv2d_*: the padded
copy ``in_fmap`` the reference builds is never materialised.
"""
import numpy as np

from .base import Layer
from .. import initializers, utils, _lib
from ..darray import DArray, LazyDArray, ParamDict
from ..regularizers import reg_coeffs
from ..runtime import config
from .._lib import Reg, Win


class Conv(Layer):
    """Constructor logic of convolution.py:7-71."""

    def __init__(self, l_in, filters, kernel_size, strides, padding, dilation_rate,
                 kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, bias_regularizer=None, name="Conv"):
        super().__init__(name=name)
        self.parents.append(l_in)

        if len(kernel_size) == len(l_in.oshape):
            input_size = l_in.oshape
            self.kernel_size = kernel_size
        else:
            input_size = l_in.oshape[1:]
            self.kernel_size = (l_in.oshape[0], *kernel_size)

        self.filters = filters
        self.strides = strides
        self.padding = padding
        self.dilation_rate = dilation_rate
        if tuple(dilation_rate) != (1,) * len(tuple(dilation_rate)):
            # Q20: the reference changes oshape but not the arithmetic — unsupported rather than wrong
            raise NotImplementedError("dilation_rate != 1 is not supported")

        _oshape = utils.get_output(input_size=input_size, kernel_size=kernel_size, strides=strides,
                                   padding=padding, dilation_rate=dilation_rate)
        _pads = utils.get_padding(padding=padding, input_size=input_size, output_size=_oshape,
                                  kernel_size=kernel_size, strides=strides)
        self.pads = tuple(_pads)
        self.oshape = tuple([self.filters] + _oshape.tolist())

        self.params = ParamDict(w1=np.zeros((self.filters, *self.kernel_size)), b1=np.zeros((self.filters,)))
        self.grads = ParamDict(w1=np.zeros((self.filters, *self.kernel_size)), b1=np.zeros((self.filters,)))

        self.kernel_initializer = kernel_initializer if kernel_initializer is not None else initializers.RandomUniform()
        # Q3 (convolution.py:52-55)
        self.bias_initializer = initializers.Zeros() if bias_initializer is None else kernel_initializer
        self.kernel_regularizer = kernel_regularizer
        self.bias_regularizer = bias_regularizer

        self.shape_pad_in = None
        self.in_fmap = None

    def initialize(self, optimizer=None):
        super().initialize(optimizer=optimizer)
        self.kernel_initializer.apply(self.params, ['w1'])
        self.bias_initializer.apply(self.params, ['b1'])

    def _win(self, batch):
        c, h, w = self.parents[0].oshape
        f, oh, ow = self.oshape
        kh, kw = self.kernel_size[-2], self.kernel_size[-1]
        if min(self.pad_top, self.pad_bottom, self.pad_left, self.pad_right) < 0:
            raise ValueError(f"{self.name}: negative padding (kernel smaller than stride with 'same')")
        return Win(batch, c, h, w, f, oh, ow, kh, kw, self.strides[0], self.strides[1],
                   self.pad_top, self.pad_bottom, self.pad_left, self.pad_right)


class Conv2D(Conv):
    """convolution.py:175-273 (the live, second definition — Q21)."""

    def __init__(self, l_in, filters, kernel_size, strides=(1, 1), padding="same", dilation_rate=(1, 1),
                 kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, bias_regularizer=None, name="Conv2D"):
        super().__init__(l_in, filters=filters, kernel_size=kernel_size, strides=strides, padding=padding,
                         dilation_rate=dilation_rate, kernel_initializer=kernel_initializer,
                         bias_initializer=bias_initializer, kernel_regularizer=kernel_regularizer,
                         bias_regularizer=bias_regularizer, name=name)
        (self.pad_top, self.pad_bottom), (self.pad_left, self.pad_right) = utils.get_side_paddings(self.pads)
        assert len(self.kernel_size) == 3
        assert len(self.strides) == 2

    # ---- fused Conv2D -> MaxPool2D -> LeakyRelu/Relu block (csrc/fused_cpa.cu), planned by Net.build -----------------
    _cpa = None           # (pool, act) when this layer heads a fusable block
    _cpa_live = None      # state of the block between this step's forward and backward

    def _cpa_usable(self):
        return (self._cpa is not None and not self.debug
                and self.kernel_regularizer is None and self.bias_regularizer is None)

    def _forward_fused(self, x):
        pool, act = self._cpa
        b = x.shape[0]
        g, gp = self._win(b), pool._win(b)
        n_pool = b * int(np.prod(pool.oshape))
        aout = act._buf('out', (b, *act.oshape))
        ig = self._raw('cpa_idx_gate', n_pool)                  # one byte per pooled element: argmax offset | gate << 4
        tc_ok = self.__dict__.setdefault("_cpa_tc_ok", {})
        if config.math not in tc_ok:        # BF16 math mode: the block's forward on the tensor core (csrc/tc_conv_c1p.cu)
            tc_ok[config.math] = bool(_lib.load().dnb_conv_pool_act_tc_supported(self._win(1), pool._win(1), config.math))
        if tc_ok[config.math]:
            self._call("dnb_conv_pool_act_fwd_tc", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), aout.ptr(), ig,
                       g, gp, float(act.alpha), config.math)
        else:
            self._call("dnb_conv_pool_act_fwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), aout.ptr(), ig,
                       g, gp, float(act.alpha))
        # the tensors the fused kernel never wrote stay observable (computed by the layer-by-layer kernels on a read)
        w_now, b_now = self._snapshot('w_snap_f', self.params['w1']), self._snapshot('b_snap_f', self.params['b1'])
        math = config.math

        def conv_out():
            out = self._buf('out', (b, *self.oshape))
            ws = self._raw('ws', _lib.load().dnb_conv2d_ws_bytes(g, math))
            self._call("dnb_conv2d_fwd", x.ptr(), w_now.ptr(), b_now.ptr(), out.ptr(), g, ws, math)
            return out.dev()
        self.output = LazyDArray((b, *self.oshape), conv_out)
        conv_lazy = self.output

        def pool_out():
            out = pool._buf('out', (b, *pool.oshape))
            idx = pool._buf('idx', (b, *pool.oshape), is_int=True)
            pool._call("dnb_maxpool2d_fwd", conv_lazy.ptr(), out.ptr(), idx.ptr(), gp)
            return out.dev()

        def pool_idx():
            idx = pool._buf('idx_unpacked', (b, *pool.oshape), is_int=True)
            pool._call("dnb_idx_gate_unpack", ig, idx.ptr(), n_pool)
            return idx.dev()
        pool.output = LazyDArray((b, *pool.oshape), pool_out)
        pool.output_idxs = LazyDArray((b, *pool.oshape), pool_idx, is_int=True)
        act.output = aout
        pool._cpa_skip_fwd = act._cpa_skip_fwd = True
        self._cpa_live = dict(ig=ig, idx=pool.output_idxs, x=x)

    # ---- tensor-core Conv2D whose epilogue runs the MaxPool2D -> LeakyRelu/Relu pair behind it (csrc/tc_conv_sv.cu) ------
    _tcp = None           # the pool, when it is this layer's only consumer and heads a fused pair (planned by Net.build)

    def _tcp_usable(self, b):
        pool = self._tcp
        if pool is None or self.debug or config.math < _lib.MATH_TF32 or not pool._pa_usable(b):
            return False
        ok = self.__dict__.setdefault("_tcp_ok", {})
        if config.math not in ok:
            ok[config.math] = bool(_lib.load().dnb_conv2d_pool_act_supported(self._win(1), pool._win(1), config.math))
        return ok[config.math]

    def _forward_tc_pool(self, x):
        """convolution + pool + activation in ONE launch: only the activation's output and one byte per pooled element are
        written; this layer's output stays observable (the plain kernel runs if somebody reads it)."""
        pool = self._tcp
        b = x.shape[0]
        g, gp, math = self._win(b), pool._win(b), config.math
        aout, ig = pool._pair_buffers(b)
        ws = self._raw('ws', _lib.load().dnb_conv2d_ws_bytes(g, math))
        self._call("dnb_conv2d_pool_act_fwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), aout.ptr(), ig,
                   g, gp, float(pool._pa.alpha), ws, math)
        w_now, b_now = self._snapshot('w_snap_f', self.params['w1']), self._snapshot('b_snap_f', self.params['b1'])

        def conv_out():
            out = self._buf('out', (b, *self.oshape))
            self._call("dnb_conv2d_fwd", x.ptr(), w_now.ptr(), b_now.ptr(), out.ptr(), g, ws, math)
            return out.dev()
        self.output = LazyDArray((b, *self.oshape), conv_out)
        pool._install_pair(self.output, aout, ig, b)
        pool._pair_done = True

    def forward(self):
        x = self._in()
        self._cpa_live = None
        if self._cpa_usable():
            return self._forward_fused(x)
        b = x.shape[0]
        if self._tcp_usable(b):
            return self._forward_tc_pool(x)
        g = self._win(b)
        out = self._buf('out', (b, *self.oshape))
        ws = self._raw('ws', _lib.load().dnb_conv2d_ws_bytes(g, config.math))
        self._call("dnb_conv2d_fwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), out.ptr(),
                   g, ws, config.math)
        self.output = out

    def _backward_fused(self, live):
        """gW / gb straight from the activation's delta (either pool routing); the parent's delta stays lazy."""
        pool, act = self._cpa
        x = live['x']
        b = x.shape[0]
        g, gp = self._win(b), pool._win(b)
        if pool.output_idxs is not live['idx']:        # the caller assigned its own argmax tensor: route through it
            user = DArray.wrap(pool.output_idxs, is_int=True)
            self._call("dnb_idx_gate_set", user.ptr(), live['ig'], user.size)
        if config.route == _lib.ROUTE_PER_SAMPLE or b == 1:
            self._call("dnb_conv_pool_act_bwd", x.ptr(), act._delta().ptr(), live['ig'], self.grads['w1'].ptr(),
                       self.grads['b1'].ptr(), g, gp, float(act.alpha), self._inv_m(b))
        else:       # pooling.py:95 as written: only sample 0 receives a delta, the last sample per argmax offset wins
            ws = self._raw('cpa_last', _lib.load().dnb_conv_pool_act_bwd_literal_ws_bytes(g, gp))
            self._call("dnb_conv_pool_act_bwd_literal", x.ptr(), act._delta().ptr(), live['ig'], self.grads['w1'].ptr(),
                       self.grads['b1'].ptr(), g, gp, float(act.alpha), self._inv_m(b), ws)
        self.grads['w1'].written()
        self.grads['b1'].written()
        w_now, math = self._snapshot('w_snap', self.params['w1']), config.math

        def materialise():              # Input.delta on demand: pool.delta -> conv.delta -> data gradient, layer by layer
            dy = self._delta()
            out = self._buf('dx', x.shape)
            ws = self._raw('ws', _lib.load().dnb_conv2d_ws_bytes(g, math))
            self._call("dnb_conv2d_bwd_data", w_now.ptr(), dy.ptr(), out.ptr(), g, ws, math)
            return out.dev()
        self.parents[0].delta = LazyDArray(x.shape, materialise)

    def backward(self):
        live, self._cpa_live = self._cpa_live, None
        if live is not None and self._cpa_usable() and self._lazy_dx:
            return self._backward_fused(live)
        x, dy = self._in(), self._delta()
        b = x.shape[0]
        g = self._win(b)
        lazy = self._lazy_dx and not self.debug
        dx = None if lazy else self._buf('dx', x.shape)
        ws = self._raw('ws', _lib.load().dnb_conv2d_ws_bytes(g, config.math))
        self._call("dnb_conv2d_bwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), dy.ptr(),
                   None if lazy else dx.ptr(), self.grads['w1'].ptr(), self.grads['b1'].ptr(), g, self._inv_m(b),
                   Reg(*reg_coeffs(self.kernel_regularizer)), Reg(*reg_coeffs(self.bias_regularizer)),
                   config.reg_scale, ws, config.math)
        self.grads['w1'].written()
        self.grads['b1'].written()
        if lazy:
            # the parent is an Input: its delta is dead on the training path (12% of the MNIST-CNN step).  Computed from
            # this step's dy buffer and a snapshot of the weights only if somebody reads `l_in.delta`.
            w_now, math = self._snapshot('w_snap', self.params['w1']), config.math

            def materialise():
                out = self._buf('dx', x.shape)
                self._call("dnb_conv2d_bwd_data", w_now.ptr(), dy.ptr(), out.ptr(), g, ws, math)
                return out.dev()
            dx = LazyDArray(x.shape, materialise)
        self.parents[0].delta = dx


class PointwiseConv2D(Conv2D):
    """convolution.py:276-284: 1x1, stride 1, no padding."""

    def __init__(self, l_in, filters, kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, bias_regularizer=None, name="PointwiseConv2D"):
        super().__init__(l_in, filters, kernel_size=(1, 1), strides=(1, 1), dilation_rate=(1, 1), padding="none",
                         kernel_initializer=kernel_initializer, bias_initializer=bias_initializer,
                         kernel_regularizer=kernel_regularizer, bias_regularizer=bias_regularizer, name=name)


class DepthwiseConv2D(Conv):
    """convolution.py:287-388; ``depth_multiplier`` is stored but unused, as in the reference."""

    def __init__(self, l_in, kernel_size, strides=(1, 1), padding="same", dilation_rate=(1, 1), depth_multiplier=1,
                 kernel_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, bias_regularizer=None, name="DepthwiseConv2D"):
        if len(l_in.oshape) != 3:
            raise ValueError(f"Expected a 3D layer ({name})")
        filters = l_in.oshape[0]
        self.depth_multiplier = int(depth_multiplier)
        super().__init__(l_in, filters=filters, kernel_size=kernel_size, strides=strides, padding=padding,
                         dilation_rate=dilation_rate, kernel_initializer=kernel_initializer,
                         bias_initializer=bias_initializer, kernel_regularizer=kernel_regularizer,
                         bias_regularizer=bias_regularizer, name=name)
        (self.pad_top, self.pad_bottom), (self.pad_left, self.pad_right) = utils.get_side_paddings(self.pads)
        assert len(self.kernel_size) == 3
        assert len(self.strides) == 2
        self.kernel_size = (1, *list(self.kernel_size)[1:])
        self.params = ParamDict(w1=np.zeros((self.filters, *self.kernel_size)), b1=np.zeros((self.filters,)))
        self.grads = ParamDict(w1=np.zeros((self.filters, *self.kernel_size)), b1=np.zeros((self.filters,)))

    def forward(self):
        x = self._in()
        b = x.shape[0]
        out = self._buf('out', (b, *self.oshape))
        self._call("dnb_dwconv2d_fwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), out.ptr(),
                   self._win(b))
        self.output = out

    def backward(self):
        if self.kernel_regularizer:
            # the reference raises here too (shape bug at convolution.py:380)
            raise ValueError("non-broadcastable output operand: DepthwiseConv2D does not support a kernel_regularizer")
        x, dy = self._in(), self._delta()
        b = x.shape[0]
        dx = self._buf('dx', x.shape)
        self._call("dnb_dwconv2d_bwd", x.ptr(), self.params['w1'].ptr(), self.params['b1'].ptr(), dy.ptr(),
                   dx.ptr(), self.grads['w1'].ptr(), self.grads['b1'].ptr(), self._win(b), self._inv_m(b),
                   Reg(*reg_coeffs(self.bias_regularizer)), config.reg_scale)
        self.grads['w1'].written()
        self.grads['b1'].written()
        self.parents[0].delta = dx
