"""SimpleRNN (dnpy/layers/recurrent.py:31-266): non-stateful, last-output recurrent layer whose
timestep loop (and the truncated-BPTT inner loop) runs inside one persistent CUDA kernel."""
import numpy as np

from .base import Layer
from .. import _lib
from ..runtime import config
from .. import initializers
from ..darray import ParamDict


class RNNCell:
    """recurrent.py:7-29 — kept for API compatibility; the arithmetic lives in dnb_rnn_fwd/bwd."""

    def __init__(self, params, grads, name="RNNCell"):
        self.params = params
        self.grads = grads


class BaseRNN(Layer):
    """recurrent.py:31-99."""

    def __init__(self, l_in, hidden_dim, output_dim, cell, params, grads,
                 kernel_initializer=None, recurrent_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, recurrent_regularizer=None, bias_regularizer=None,
                 return_sequences=False, return_state=False,
                 stateful=False, unroll=False, bptt_truncate=None, name="BaseRNN"):
        super().__init__(name=name)
        self.parents.append(l_in)
        self.states_h = []
        self.sequences = []
        self.hidden_dim = hidden_dim
        self.output_dim = output_dim
        self.params = params
        self.grads = grads
        self.cell = cell
        self.stateful = stateful
        self.unroll = unroll
        self.return_sequences = return_sequences
        self.return_state = return_state
        self.bptt_truncate = bptt_truncate
        self.oshape = (None, self.output_dim) if self.return_sequences else (self.output_dim,)
        self.kernel_initializer = kernel_initializer if kernel_initializer is not None else initializers.RandomUniform()
        self.recurrent_initializer = recurrent_initializer if recurrent_initializer is not None \
            else initializers.RandomUniform()
        # Q3 (recurrent.py:71-74)
        self.bias_initializer = initializers.Zeros() if bias_initializer is None else kernel_initializer
        self.kernel_regularizer = kernel_regularizer
        self.recurrent_regularizer = recurrent_regularizer
        self.bias_regularizer = bias_regularizer
        self.sublayers = []
        if unroll:
            raise NotImplementedError("Unrolled layer not yet implemented")
        if stateful or return_sequences:
            # the reference's stateful branch is `pass`; return_sequences has no delta path here
            raise NotImplementedError("stateful / return_sequences RNNs are out of scope (SURVEY.md §2 row 9)")


class SimpleRNN(BaseRNN):
    """recurrent.py:102-266."""

    def __init__(self, l_in, hidden_dim, output_dim=None,
                 kernel_initializer=None, recurrent_initializer=None, bias_initializer=None,
                 kernel_regularizer=None, recurrent_regularizer=None, bias_regularizer=None,
                 return_sequences=False, return_state=False,
                 stateful=False, unroll=False, bptt_truncate=None, name="SimpleRNN"):
        output_dim = output_dim if output_dim else hidden_dim
        timesteps, features = l_in.oshape
        params = ParamDict(waa=np.zeros((hidden_dim, hidden_dim)), wax=np.zeros((hidden_dim, features)),
                           ba=np.zeros((1, hidden_dim)))
        grads = ParamDict(waa=np.zeros((hidden_dim, hidden_dim)), wax=np.zeros((hidden_dim, features)),
                          ba=np.zeros((1, hidden_dim)))
        self.outputs, self.states_h = [], []
        self.delta_h_t = None
        super().__init__(l_in=l_in, hidden_dim=hidden_dim, output_dim=output_dim, cell=RNNCell,
                         params=params, grads=grads,
                         kernel_initializer=kernel_initializer, recurrent_initializer=recurrent_initializer,
                         bias_initializer=bias_initializer, kernel_regularizer=kernel_regularizer,
                         recurrent_regularizer=recurrent_regularizer, bias_regularizer=bias_regularizer,
                         return_sequences=return_sequences, return_state=return_state,
                         stateful=stateful, unroll=unroll, bptt_truncate=bptt_truncate, name=name)

    def initialize(self, optimizer=None):
        super().initialize(optimizer=optimizer)
        self.recurrent_initializer.apply(self.params, ['waa'])
        self.kernel_initializer.apply(self.params, ['wax'])
        self.bias_initializer.apply(self.params, ['ba'])

    def forward(self):
        x = self._in()
        b, t, d = x.shape
        h = self.hidden_dim
        states = self._buf('states', (b, t, h))
        out = self._buf('out', (b, h))
        self._call("dnb_rnn_fwd", x.ptr(), self.params['waa'].ptr(), self.params['wax'].ptr(), self.params['ba'].ptr(),
                   states.ptr(), out.ptr(), b, t, d, h)
        self.states_h = states
        self.outputs = states
        self.output = out

    def backward(self):
        x, dy = self._in(), self._delta()
        b, t, d = x.shape
        h = self.hidden_dim
        trunc = self.bptt_truncate if self.bptt_truncate else t
        dx = self._buf('dx', (b, t, d))
        lib = _lib.load()
        if config.math >= _lib.MATH_TF32 and lib.dnb_rnn_bwd_tc_supported(t, d, h, int(trunc), config.math):
            # tensor-core math modes: serial chain per sample + the lag recurrences of all rows as tcgen05 tiles (csrc/rnn_tc.cu)
            ws = self._raw('ws_pre', lib.dnb_rnn_bwd_tc_ws_bytes(b, t, h))
            self._call("dnb_rnn_bwd_tc", x.ptr(), self.states_h.ptr(), dy.ptr(), self.params['waa'].ptr(),
                       self.params['wax'].ptr(), dx.ptr(), self.grads['waa'].ptr(), self.grads['wax'].ptr(),
                       self.grads['ba'].ptr(), b, t, d, h, int(trunc), ws, config.math)
        else:
            self._call("dnb_rnn_bwd", x.ptr(), self.states_h.ptr(), dy.ptr(), self.params['waa'].ptr(),
                       self.params['wax'].ptr(), dx.ptr(), self.grads['waa'].ptr(), self.grads['wax'].ptr(),
                       self.grads['ba'].ptr(), b, t, d, h, int(trunc))
        for k in ('waa', 'wax', 'ba'):
            self.grads[k].written()
        self.x_t_deltas = dx
        self.parents[0].delta = dx
