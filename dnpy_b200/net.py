"""Net — the orchestrator with dnpy's API (dnpy/net.py) driving CUDA kernels.

Public surface and print strings are the reference's: ``build / fit / evaluate /
set_mode / feed_input / forward / compute_losses / compute_metrics / reset_grads /
do_delta / backward / apply_grads / summary / get_params / set_params / get_grads /
set_grads / save / load``.  What changes underneath:

* after ``build`` the trainable parameters, their gradients and the optimizer state
  live in FLAT fp32 arenas (one allocation each), so ``reset_grads`` is one memset,
  ``apply_grads`` is one fused optimizer launch for the whole net, and the
  data-parallel exchange is one all-reduce on the gradient arena;
* ``fit`` can replay a captured CUDA graph of the whole step (forward -> loss ->
  reset -> backward -> [all-reduce] -> optimizer) instead of ~40 Python-driven launches;
* with ``dnpy_b200.parallel`` initialised (one process per GPU), ``fit`` shards every
  minibatch over the ranks.
"""
import copy
import math
import os
import pickle

import numpy as np

from . import _lib
from .darray import DArray, ParamDict, dev_empty, dev_zeros, require_cuda, stream_ptr, torch
from .layers import Softmax
from .losses import CrossEntropy
from .optimizers import Adam, RMSProp, SGD
from .runtime import config

_ALIGN = 32     # floats: every arena slot starts on a 128-byte boundary (TMA / float4 friendly)


class _ResidentSlice:
    """A row block of a dataset that Net.fit uploaded once: a CUDA tensor view.  `_stage_in` copies it
    device-to-device into the step's persistent input buffer, so buffer addresses (and the captured step graph) do not
    depend on the minibatch index."""
    __slots__ = ("t",)

    def __init__(self, t):
        self.t = t

    def __len__(self):
        return int(self.t.shape[0])

    def __getitem__(self, key):
        return _ResidentSlice(self.t[key])

    @property
    def shape(self):
        return tuple(self.t.shape)


def check_datasets(x_train, y_train, x_test, y_test):
    """net.py:10-35."""
    if x_train and y_train and x_test and y_test:
        assert len(x_train) == len(x_test)
        assert len(y_train) == len(y_test)
        for a, b in zip(x_train, x_test):
            assert a.shape[1:] == b.shape[1:]
        for a, b in zip(y_train, y_test):
            assert a.shape[1:] == b.shape[1:]
    if x_train and y_train:
        for a, b in zip(x_train, y_train):
            assert len(a) == len(b)
    if x_test and y_test:
        for a, b in zip(x_test, y_test):
            assert len(a) == len(b)


def _visit(node, temp, done, order):
    """Depth-first post-order with cycle detection (net.py:38-58): parents first, then the node is
    pushed to the FRONT, so ``order`` ends up outputs-first."""
    key = id(node)
    if key in done:
        return
    if key in temp:
        raise RecursionError("Not a DAG")
    temp.add(key)
    for parent in node.parents:
        _visit(parent, temp, done, order)
    temp.remove(key)
    done.add(key)
    order.insert(0, node)


class Net:

    def __init__(self):
        self.l_in = None
        self.l_out = None
        self.fts_layers = None
        self.bts_layers = None
        self.optimizer = None
        self.losses = None
        self.metrics = None
        self.debug = False
        self.mode = 'test'
        self.smart_derivatives = None
        # device state
        self._arena = None
        self.use_graph = os.environ.get("DNPY_B200_GRAPH", "1") != "0"
        self._graphs = {}
        self._pending_flags = []
        self._stage = {}
        self._copy_stream = None
        self.graph_launches = 0          # kernels executed through graph replays (the library counts eager ones)

    # ------------------------------------------------------------------ build
    def build(self, l_in, l_out, optimizer, losses, metrics, debug=False, smart_derivatives=False):
        self.l_in = l_in
        self.l_out = l_out
        self.fts_layers = []
        self.bts_layers = []
        self.optimizer = optimizer
        self.losses = losses
        self.metrics = metrics
        self.debug = debug
        self.smart_derivatives = smart_derivatives
        self.bts()
        self.fts()
        self.initialize()
        self._arena = None
        self._graphs = {}
        self._plan_fusions()

    def _plan_fusions(self):
        """Decided once per build: dead data gradients (lazy), the fused first block Conv2D(C=1) -> MaxPool2D -> LeakyRelu/Relu
        (csrc/fused_cpa.cu) and fused MaxPool2D -> LeakyRelu/Relu pairs (dnb_maxpool2d_act_fwd / dnb_pool_act_bwd).  Every
        layer still exposes its own output / delta: what a fused kernel never wrote is a LazyDArray."""
        import os
        from .layers import MaxPool2D, LeakyRelu, Conv2D, Dense, Input, Reshape, DepthwiseConv2D
        # dead data gradients: a Conv2D / Dense whose delta would only travel (through Reshape views) to an Input layer
        # skips it; the Input's `delta` becomes a LazyDArray that runs the data-gradient kernel if somebody reads it
        for l in self.fts_layers:
            l._lazy_dx = False
            if isinstance(l, (Conv2D, Dense)) and not isinstance(l, DepthwiseConv2D) and not os.environ.get("DNPY_B200_EAGER_DX"):
                p = l.parents[0]
                while isinstance(p, Reshape):
                    p = p.parents[0]
                l._lazy_dx = isinstance(p, Input)
        consumers = {}
        for l in self.fts_layers:
            for p in l.parents:
                consumers.setdefault(id(p), []).append(l)
        # fused first block (csrc/fused_cpa.cu): Conv2D(C=1, 3x3) -> MaxPool2D(3x3/2) -> LeakyRelu/Relu, each the only
        # consumer of its parent: one forward kernel and one backward kernel, the conv output / gradient never reach HBM
        for l in self.fts_layers:
            l._cpa = None
            l._cpa_block = None
        for l in self.fts_layers:
            cons = consumers.get(id(l), [])
            if type(l) is not Conv2D or len(cons) != 1 or type(cons[0]) is not MaxPool2D or os.environ.get("DNPY_B200_NO_CPA"):
                continue
            pool = cons[0]
            cons2 = consumers.get(id(pool), [])
            if len(cons2) != 1 or type(cons2[0]).__name__ not in ("LeakyRelu", "Relu") or not isinstance(cons2[0], LeakyRelu):
                continue
            act = cons2[0]
            if l in self.l_out or pool in self.l_out or not l._lazy_dx:
                continue
            try:
                ok = _lib.load().dnb_conv_pool_act_supported(l._win(1), pool._win(1))
            except ValueError:
                ok = 0
            if ok:
                l._cpa = (pool, act)
                pool._cpa_block = act._cpa_block = l
        # MaxPool2D -> LeakyRelu/Relu pairs outside a fused first block: the pool's kernel writes the activation's output and
        # one byte per pooled element; the backward pair is one pass from the activation's delta
        for l in self.fts_layers:
            if isinstance(l, MaxPool2D):
                l._pa = None
        for l in self.fts_layers:
            cons = consumers.get(id(l), [])
            if (not isinstance(l, MaxPool2D) or len(cons) != 1 or type(cons[0]).__name__ not in ("LeakyRelu", "Relu")
                    or not isinstance(cons[0], LeakyRelu) or l in self.l_out or l._cpa_block is not None
                    or os.environ.get("DNPY_B200_NO_POOL_ACT")):
                continue
            try:
                ok = _lib.load().dnb_pool_act_supported(l._win(1))
            except ValueError:
                ok = 0
            if ok:
                l._pa = cons[0]
                cons[0]._cpa_block = l
        # a tensor-core Conv2D whose only consumer heads such a pair runs the pair in its epilogue (decided per step:
        # tf32 / bf16 math modes, geometry taken by dnb_conv2d_pool_act_fwd)
        for l in self.fts_layers:
            if type(l) is Conv2D:
                cons = consumers.get(id(l), [])
                l._tcp = None
                if (len(cons) == 1 and isinstance(cons[0], MaxPool2D) and cons[0]._pa is not None and l not in self.l_out
                        and not os.environ.get("DNPY_B200_NO_CPA_TC")):
                    l._tcp = cons[0]

    def fts(self):
        self.fts_layers = list(self.bts_layers)
        self.fts_layers.reverse()

    def bts(self):
        self.bts_layers = []
        temp, done = set(), set()
        for node in self.l_out:
            if id(node) not in done:
                _visit(node, temp, done, self.bts_layers)
        assert not temp

    def initialize(self):
        """net.py:277-293: initialise in forward order (this fixes the order in which the global
        NumPy stream is consumed, hence the weights), then wire the CE(Softmax) shortcut."""
        batch_size = self.fts_layers[0].batch_size
        for l in self.fts_layers:
            l.debug = self.debug
            l.initialize(optimizer=self.optimizer)
            l.batch_size = batch_size
        if self.smart_derivatives:
            for l in self.l_out:
                for lo in self.losses:           # all (output, loss) pairs, like the reference (Q19)
                    if isinstance(l, Softmax) and isinstance(lo, CrossEntropy):
                        print("[INFO] Using derivative speed-up: 'CrossEntropy(Softmax(z))'")
                        lo.softmax_output = True
                        l.ce_loss = True

    # ------------------------------------------------------------------ arenas
    def _trainable(self):
        out = []
        for l in self.bts_layers:
            for k in l.params.keys():
                if k in l.grads:
                    out.append((l, k))
        return out

    def _bind_arenas(self):
        """Lay every trainable tensor out in flat fp32 arenas: params P, grads G and the optimizer
        state (V and/or S).  Idempotent; re-run after ``set_params`` replaced the dicts."""
        require_cuda()
        slots = self._trainable()
        if self._arena is not None and self._arena["keys"] == [(id(l.params[k]), id(l.grads[k])) for l, k in slots]:
            return self._arena
        t = torch()
        offs, pos = [], 0
        for l, k in slots:
            offs.append(pos)
            pos += (l.params[k].size + _ALIGN - 1) // _ALIGN * _ALIGN
        total = max(pos, _ALIGN)
        P, G = dev_zeros((total,)), dev_zeros((total,))
        peer = None
        if self._arena is not None and self._arena.get("peer") is not None:
            self._arena["peer"].close()             # the arena is being rebuilt (set_params replaced the dicts)
        if config.world > 1:
            # data parallel on one node: the gradient arena lives in CUDA-IPC memory every rank maps; the exchange is one kernel
            # over NVLink peer memory (parallel.PeerArena, csrc/dp_peer.cu) instead of NCCL all-reduces
            from . import parallel
            import torch.distributed as dist
            if dist.is_initialized() and dist.get_backend() == "nccl" and parallel.peer_exchange_possible():
                peer = parallel.PeerArena(total)
                G = peer.tensor[:total]
        states = {}
        kind = type(self.optimizer).__name__
        for name in ("V", "S"):
            if hasattr(self.optimizer, name):
                states[name] = dev_zeros((total,))
        uniform = True
        for (l, k), o in zip(slots, offs):
            n = l.params[k].size
            shape = l.params[k].shape
            l.params[k].bind(P[o:o + n].view(*shape))
            l.grads[k].bind(G[o:o + n].view(*shape))
            opt = l.optimizer
            if opt is None or type(opt).__name__ != kind:
                uniform = False
                continue
            for name, arena in states.items():
                st = getattr(opt, name)
                if k in st:
                    st[k].bind(arena[o:o + n].view(*shape))
        self._arena = dict(P=P, G=G, states=states, total=total, uniform=uniform, offsets=offs, peer=peer,
                           keys=[(id(l.params[k]), id(l.grads[k])) for l, k in slots], slots=slots)
        self._graphs = {}
        return self._arena

    # ------------------------------------------------------------------ step primitives (net.py:295-364)
    def do_delta(self, deltas):
        for i in range(len(self.l_out)):
            self.l_out[i].delta = deltas[i]

    def reset_grads(self):
        if self._device_ready():
            a = self._bind_arenas()
            _lib.call("dnb_fill_f32", a["G"].data_ptr(), a["total"], 0.0, stream_ptr())
            for l, k in a["slots"]:
                l.grads[k].written()
            bound = {id(l.grads[k]) for l, k in a["slots"]}
            for l in self.fts_layers:
                for k in l.grads.keys():
                    if id(l.grads[k]) not in bound:
                        l.grads[k].fill(0.0)
            return
        for l in self.fts_layers:
            for k in l.grads.keys():
                l.grads[k].fill(0.0)

    def feed_input(self, x, key="x"):
        """net.py:304-307.  On a CUDA device this is the H2D boundary: each input minibatch is copied
        (asynchronously when the source is pinned) into a persistent device buffer, so its address
        is stable across steps."""
        ready = self._device_ready()
        for j in range(len(self.l_in)):
            self.l_in[j].output = self._stage_in((key, j), x[j], self._input_is_int(j)) if ready else x[j]

    def _input_is_int(self, j):
        from .layers import Embedding
        return any(isinstance(l, Embedding) and l.parents and l.parents[0] is self.l_in[j] for l in self.fts_layers)

    def _stage_in(self, key, a, is_int=False):
        """ndarray / torch CPU tensor -> persistent device buffer (fp32, or int32 for token ids).

        Two buffers per input alternate between steps and PINNED sources are copied on a dedicated copy
        stream, so the H2D transfer of step k overlaps the kernels of step k-1 (the compute stream only
        waits for the copy's event; a buffer is reused once the step that read it has drained)."""
        if isinstance(a, DArray):
            return a
        t = torch()
        resident = isinstance(a, _ResidentSlice)
        if resident:
            src = a.t
        elif t.is_tensor(a):
            src = a
            if src.is_cuda:
                return DArray.device(src, is_int=is_int)
        else:
            arr = np.asarray(a)
            want = np.int32 if is_int else np.float32
            if arr.dtype != want or not arr.flags.c_contiguous:
                arr = np.ascontiguousarray(arr, dtype=want)
            src = t.from_numpy(arr)
        want_t = t.int32 if is_int else t.float32
        if src.dtype != want_t:
            src = src.to(want_t)
        st = self._stage.get(key)
        if st is None or tuple(st["bufs"][0].shape) != tuple(src.shape):
            st = dict(bufs=[dev_empty(tuple(src.shape), int_dtype=is_int) for _ in range(2)],
                      free=[t.cuda.Event(), t.cuda.Event()], ready=[t.cuda.Event(), t.cuda.Event()], k=0)
            self._stage[key] = st
        cur = t.cuda.current_stream()
        k = st["k"]
        st["k"] = k + 1
        slot, prev = k & 1, (k - 1) & 1
        if k > 0:
            st["free"][prev].record(cur)          # everything queued so far (the previous step) has read slot `prev`
        buf = st["bufs"][slot]
        if resident:
            buf.copy_(src, non_blocking=True)       # device-to-device, stream ordered behind the previous step
        elif src.is_pinned():
            if self._copy_stream is None:
                self._copy_stream = t.cuda.Stream()
            cs = self._copy_stream
            if k > 1:
                cs.wait_event(st["free"][slot])   # recorded one call ago: the step that last read this slot is covered
            with t.cuda.stream(cs):
                buf.copy_(src, non_blocking=True)
                st["ready"][slot].record(cs)
            cur.wait_event(st["ready"][slot])
        else:
            buf.copy_(src, non_blocking=True)       # pageable source: stream-ordered staging copy
        return DArray.device(buf, is_int=is_int)

    def forward(self):
        for l in self.fts_layers:
            l.forward()
            if self.debug:
                l.print_stats()

    def backward(self):
        buckets = peer = None
        if config.world > 1:
            from . import parallel
            peer = self._bind_arenas()["peer"]
            if peer is None:
                buckets = parallel.GradBuckets(self)
        for l in self.bts_layers:
            l.backward()
            if self.debug:
                l.print_stats()
            if buckets is not None:
                buckets.layer_done(l)          # this layer's gradients are final: start their all-reduce now
        if peer is not None:
            peer.allreduce()                   # one kernel: every rank's arena now holds the global sum
        if buckets is not None:
            buckets.finish()
        if peer is not None or buckets is not None:
            for l, k in self._arena["slots"]:
                l.grads[k].written()

    @staticmethod
    def _opt_hyper(o):
        """Class + scalar hyper-parameters of an optimizer copy (state dicts excluded)."""
        if o is None:
            return None
        return (type(o).__name__,) + tuple(sorted((k, v) for k, v in vars(o).items()
                                                 if isinstance(v, (int, float, bool, str)) and k != "name"))

    def _uniform_hyper(self):
        """The hyper-parameters shared by EVERY layer's deep-copied optimizer (base.py:34), or None when they differ
        (layer-wise lr, a user poking ``l.optimizer.lr``): only then is one launch over the whole arena legal."""
        hs = {self._opt_hyper(l.optimizer) for l in self.bts_layers if any(k in l.grads for k in l.params.keys())}
        return next(iter(hs)) if len(hs) == 1 else None

    def apply_grads(self):
        a = self._bind_arenas() if self._device_ready() else None
        if a is not None and a["uniform"] and not any(l.frozen for l in self.bts_layers) \
                and self._uniform_hyper() is not None:
            self._apply_fused(a)
            return
        for l in self.bts_layers:
            if not l.frozen:
                l.optimizer.apply(l.params, l.grads)

    def _apply_fused(self, a):
        """One optimizer launch over the whole arena: legal because every layer's optimizer copy
        carries identical hyper-parameters (base.py:34)."""
        o = a["slots"][0][0].optimizer          # the layers' copies carry the live hyper-parameters, like the reference
        p, g, n = a["P"].data_ptr(), a["G"].data_ptr(), a["total"]
        if isinstance(o, Adam):
            if o.bias_correction:
                raise TypeError("unsupported operand type(s) for ** or pow(): 'float' and 'NoneType'")
            o.launch(p, g, a["states"]["V"].data_ptr(), a["states"]["S"].data_ptr(), n)
        elif isinstance(o, RMSProp):
            o.launch(p, g, a["states"]["S"].data_ptr(), n)
        elif isinstance(o, SGD):
            if o.bias_correction and o.momentum != 0.0:
                raise TypeError("unsupported operand type(s) for ** or pow(): 'float' and 'NoneType'")
            o.launch(p, g, a["states"]["V"].data_ptr(), n)
        else:
            raise NotImplementedError(type(o).__name__)
        for l, k in a["slots"]:
            l.params[k].written()

    def compute_losses(self, y_target):
        losses, deltas = [], []
        for i in range(len(self.losses)):
            y_pred_i = self.l_out[i].output
            y_target_i = y_target[i]
            loss = self.losses[i].compute_loss(y_pred_i, y_target_i)
            losses.append(loss)
            delta = self.losses[i].compute_delta(y_pred_i, y_target_i)
            deltas.append(delta)
            self._raise_on_flags(loss, self.losses[i], y_pred_i, y_target_i)
        return losses, deltas

    @staticmethod
    def _raise_on_flags(loss, loss_obj, y_pred, y_target, flags=None):
        """net.py:342-350, same messages."""
        if math.isnan(loss):
            raise ValueError("NaNs in the loss function")
        elif math.isinf(loss):
            raise ValueError("Inf in the loss function")
        if flags is None:
            # the fused loss kernel's [loss, flags] tensor of the launch that just produced loss and delta (the
            # memo itself is consumed by the compute_loss / compute_delta pair)
            out = getattr(loss_obj, "_last_out", None)
            flags = int(out[1].item()) if out is not None else 0
        if flags & 4:
            raise ValueError("NaNs in the delta")
        elif flags & 8:
            raise ValueError("Info in the delta")

    def compute_metrics(self, y_target):
        metrics = []
        for i in range(len(self.l_out)):
            row = []
            for me in self.metrics[i]:
                row.append(me.compute_metric(self.l_out[i].output, y_target[i]))
            metrics.append(row)
        return metrics

    def set_mode(self, mode="train"):
        self.mode = mode
        for l in self.fts_layers:
            if self.mode == "train":
                l.training = True
            elif self.mode == "test":
                l.training = False
            else:
                raise KeyError("Unknown mode")

    # ------------------------------------------------------------------ fused training step
    def _device_ready(self):
        t = torch()
        return t.cuda.is_available()

    def train_step(self, x_mb, y_mb):
        """One iteration of net.py:176-195 without host synchronisation.  Returns the per-loss device
        tensors ``[loss, flags]``; nothing is read back here.

        After two eager calls with the same shapes the whole step (forward -> loss -> reset -> backward ->
        optimizer, ~30-60 kernels) is captured in a CUDA graph per input staging slot and replayed; host-side
        work per step is then the input copies, the Dropout draws and one graph launch."""
        xs = [self._stage_in(("x", j), x_mb[j], self._input_is_int(j)) for j in range(len(self.l_in))]
        ys = [self._stage_in(("y", i), y_mb[i]) for i in range(len(self.losses))]
        if not (self.use_graph and not self.debug and self._graph_ok_for_world()):
            return self._step_body(xs, ys)
        key = (tuple(x.ptr() for x in xs), tuple(y.ptr() for y in ys), tuple(x.shape for x in xs), self.mode,
               config.math, config.route, self._uniform_hyper(), self._layer_generation())
        ent = self._graphs.get(key)
        if ent is None:
            ent = self._graphs[key] = dict(calls=0, graph=None, outs=None)
        if ent["graph"] is None:
            ent["calls"] += 1
            if ent["calls"] <= 2 or any(l.frozen for l in self.fts_layers):
                return self._step_body(xs, ys)              # warm-up: every buffer gets allocated here
            self._prepare_host_draws(xs)
            t = torch()
            t.cuda.synchronize()
            g = t.cuda.CUDAGraph()
            n0 = _lib.load().dnb_launch_count()
            with t.cuda.graph(g):
                outs = self._step_body(xs, ys)
            if self._layer_generation() != key[-1]:         # something allocated during capture: do not trust it
                del self._graphs[key]
                return self._step_body(xs, ys)
            ent["graph"], ent["outs"] = g, outs
            ent["launches"] = int(_lib.load().dnb_launch_count() - n0)
            # capturing records the step without running it: fall through to the replay
        else:
            self._prepare_host_draws(xs)
        ent["graph"].replay()
        self.graph_launches += ent["launches"]
        self._mark_written()
        return ent["outs"]

    @staticmethod
    def _graph_ok_for_world():
        """Data parallel: the gradient all-reduce is captured inside the step graph when the group runs on NCCL
        (capturable); other backends (gloo in the CPU/1-GPU tests) keep the eager step."""
        if config.world == 1:
            return True
        if os.environ.get("DNPY_B200_DP_GRAPH", "1") == "0":
            return False
        import torch.distributed as dist
        return dist.is_initialized() and dist.get_backend() == "nccl"

    def _step_body(self, xs, ys):
        for j in range(len(self.l_in)):
            self.l_in[j].output = xs[j]
        self.forward()
        outs, deltas = [], []
        for i, lo in enumerate(self.losses):
            outs.append(lo.result_tensor(self.l_out[i].output, ys[i]))
            deltas.append(lo.compute_delta(self.l_out[i].output, ys[i]))
        self.reset_grads()
        self.do_delta(deltas)
        self.backward()
        self.apply_grads()
        return outs

    @staticmethod
    def _layer_generation():
        from .layers.base import Layer
        return Layer.buffer_generation

    def _prepare_host_draws(self, xs):
        """Host-RNG layers (Dropout) draw before a graph capture/replay, in forward order, so the global
        NumPy stream is consumed exactly like an eager step."""
        for l in self.fts_layers:
            if hasattr(l, "prepare") and l.training:
                shape = l.parents[0].output.shape if l.parents[0].output is not None else None
                if shape is not None:
                    l.prepare(shape)

    def _mark_written(self):
        """After a replay the host copies of everything the step wrote are stale."""
        for l in self.fts_layers:
            for d in (l.output, l.delta, getattr(l, "output_idxs", None)):
                if isinstance(d, DArray):
                    d.written()
            for d in list(l.params.values()) + list(l.grads.values()):
                d.written()

    # ------------------------------------------------------------------ fit / evaluate
    def _format_eval(self, losses=None, metrics=None):
        str_l, str_m = [], []
        if losses is not None:
            assert len(self.l_out) == len(self.losses) == len(losses)
            for i in range(len(self.losses)):
                str_l += ["{}={:.5f}".format(self.losses[i].name, float(losses[i]))]
        if metrics is not None:
            assert len(self.l_out) == len(metrics)
            for i in range(len(self.l_out)):
                tmp = []
                for j in range(len(self.metrics[i])):
                    tmp += ["{}={:.5f}".format(self.metrics[i][j].name, float(metrics[i][j]))]
                str_m += ["{}: ({})".format(self.l_out[i].name, ", ".join(tmp))]
        return [str_l, str_m]

    def get_minibatch(self, x_train, y_train, batch_i, batch_size):
        """net.py:115-125, then this rank's row block when the batch is sharded."""
        x_mb, y_mb = [], []
        lo = batch_i * batch_size
        for j in range(len(self.l_in)):
            x_mb += [x_train[j][lo:lo + batch_size]]
            y_mb += [y_train[j][lo:lo + batch_size]]
        if config.world > 1:
            from . import parallel
            x_mb = [parallel.shard_rows(a) for a in x_mb]
            y_mb = [parallel.shard_rows(a) for a in y_mb]
        return x_mb, y_mb

    def fit(self, x_train, y_train, batch_size, epochs, x_test=None, y_test=None, evaluate_epoch=True, print_rate=1):
        assert isinstance(x_train, list)
        assert isinstance(y_train, list)
        assert isinstance(x_test, list) or x_test is None
        assert isinstance(y_test, list) or y_test is None
        assert len(self.l_in) == len(x_train)
        assert len(self.l_out) == len(self.losses) == len(y_train)
        check_datasets(x_train, y_train, x_test, y_test)

        batch_size = int(batch_size)
        num_samples = len(x_train[0])
        num_batches = int(num_samples / batch_size)
        verbose = config.rank == 0
        # dataset residency: upload the sets once (fp32 / int32) and slice minibatches on the device
        x_train, y_train = self._make_resident(x_train, y_train)
        if x_test is not None and y_test is not None:
            x_test, y_test = self._make_resident(x_test, y_test)

        for i in range(epochs):
            printing = (i % print_rate) == 0
            if printing and verbose:
                print(f"Epoch {i+1}/{epochs}...")
            if self.mode != "train":
                self.set_mode('train')
            for b in range(num_batches):
                x_mb, y_mb = self.get_minibatch(x_train, y_train, b, batch_size)
                if printing:
                    # reference order: forward, losses (+guards), metrics, print, then the update
                    y_mb = [DArray.device(v.t) if isinstance(v, _ResidentSlice) else v for v in y_mb]
                    self.feed_input(x_mb)
                    self.forward()
                    b_losses, b_deltas = self.compute_losses(y_target=y_mb)
                    b_metrics = self.compute_metrics(y_target=y_mb)
                    if config.world > 1:          # report the global batch: mean of the per-rank means
                        from . import parallel
                        b_losses = parallel.mean_over_ranks(b_losses)
                        b_metrics = [parallel.mean_over_ranks(m) for m in b_metrics]
                    str_eval = self._format_eval(b_losses, b_metrics)
                    if verbose:
                        print(f"\t - Batch {b+1}/{num_batches} - Losses[{', '.join(str_eval[0])}]; Metrics[{'; '.join(str_eval[1])}]")
                    self.reset_grads()
                    self.do_delta(b_deltas)
                    self.backward()
                    self.apply_grads()
                else:
                    # the step returns the losses' PERSISTENT [loss, flags] tensors: keep a copy per step, or only the
                    # newest step would ever be checked
                    self._pending_flags.append([o.clone() for o in self.train_step(x_mb, y_mb)])
                    if len(self._pending_flags) >= 64:
                        self._check_pending()
            self._check_pending()
            if evaluate_epoch and printing:
                self.set_mode('test')
                lo, me = self.evaluate(x_train, y_train, batch_size=1)
                str_eval = self._format_eval(lo, me)
                if verbose:
                    print(f"\t - Training losses[{', '.join(str_eval[0])}]; Training metrics[{'; '.join(str_eval[1])}];")
                if x_test is not None and y_test is not None:
                    lo, me = self.evaluate(x_test, y_test, batch_size=1)
                    str_eval = self._format_eval(lo, me)
                    if verbose:
                        print(f"\t - Validation losses[{', '.join(str_eval[0])}]; Validation metrics[{'; '.join(str_eval[1])}];")

    RESIDENT_BYTES = 16 << 30

    def _make_resident(self, xs, ys):
        """Upload a dataset once if it fits RESIDENT_BYTES (and a device is present); else leave it on the host."""
        import os
        if not self._device_ready() or os.environ.get("DNPY_B200_NO_RESIDENT"):
            return xs, ys
        if any(isinstance(a, (_ResidentSlice, DArray)) for a in list(xs) + list(ys)):
            return xs, ys
        t = torch()
        total = sum(int(np.prod(np.shape(a))) * 4 for a in list(xs) + list(ys))
        if total > self.RESIDENT_BYTES:
            return xs, ys

        def up(a, is_int):
            if t.is_tensor(a):
                src = a
            else:
                src = t.from_numpy(np.ascontiguousarray(np.asarray(a), dtype=np.int32 if is_int else np.float32))
            return _ResidentSlice(src.to("cuda", dtype=t.int32 if is_int else t.float32).contiguous())

        return ([up(a, self._input_is_int(j)) for j, a in enumerate(xs)], [up(a, False) for a in ys])

    def _check_pending(self):
        """Deferred NaN/Inf guard for the non-printing steps (one sync per 64 steps, not per step)."""
        pend, self._pending_flags = self._pending_flags, []
        for outs in pend:
            for o in outs:
                v = o.tolist()
                self._raise_on_flags(v[0], None, None, None, flags=int(v[1]))

    def evaluate(self, x_test, y_test, batch_size=1):
        assert isinstance(x_test, list)
        assert isinstance(y_test, list)
        assert len(self.l_in) == len(x_test)
        assert len(self.l_out) == len(self.losses) == len(y_test)
        check_datasets(None, None, x_test, y_test)
        num_samples = len(x_test[0])
        num_batches = int(num_samples / batch_size)
        assert batch_size <= num_samples
        if self.mode != "test":
            print("[WARNING] Setting net mode to 'test'")
            self.set_mode('test')
        if self._can_evaluate_in_chunks(batch_size):
            return self._evaluate_chunked(x_test, y_test, batch_size, num_batches)
        losses, metrics = [], []
        for b in range(num_batches):
            x_mb, y_mb = [], []
            lo_i = b * batch_size
            for j in range(len(self.l_in)):
                x_mb += [x_test[j][lo_i:lo_i + batch_size]]
                y_mb += [y_test[j][lo_i:lo_i + batch_size]]
            y_mb = [DArray.device(v.t) if isinstance(v, _ResidentSlice) else v for v in y_mb]
            self.feed_input(x_mb, key="x_eval")      # own staging buffers: the training step graph keeps its addresses
            self.forward()
            lo, de = self.compute_losses(y_target=y_mb)
            losses.append(lo)
            metrics.append(self.compute_metrics(y_target=y_mb))
        losses = np.array(losses)
        metrics = [np.array(row) for row in list(zip(*metrics))]
        losses = np.mean(losses, axis=0)
        metrics = [np.mean(me, axis=0) for me in metrics]
        assert len(losses) == len(self.losses)
        assert len(metrics) == len(self.losses)
        return losses, metrics

    # Net.fit evaluates the whole training and test set with batch_size=1 after every printed epoch (net.py:197-210):
    # one forward per SAMPLE.  In test mode every layer is independent per sample (BatchNorm uses its moving statistics,
    # Dropout scales), and MSE / BCE / CE / MAE and the accuracies are plain means over samples, so the mean of the
    # per-batch values over the first num_batches*batch_size samples equals the sample-weighted mean over ANY partition
    # of them: evaluate in chunks of up to EVAL_CHUNK samples.  Anything else (RMSE: a square root of a
    # mean) keeps the literal per-batch loop.  DNPY_B200_EVAL_LITERAL=1 forces the literal loop.
    EVAL_CHUNK = 4096
    _ADDITIVE = ("MSE", "MAE", "NLL", "Hinge", "BinaryCrossEntropy", "CrossEntropy", "BinaryAccuracy", "CategoricalAccuracy")

    def _can_evaluate_in_chunks(self, batch_size):
        import os
        if os.environ.get("DNPY_B200_EVAL_LITERAL") or batch_size >= self.EVAL_CHUNK or self.mode != "test":
            return False
        names = [type(lo).__name__ for lo in self.losses] + [type(me).__name__ for row in self.metrics for me in row]
        return all(n in self._ADDITIVE for n in names)

    def _evaluate_chunked(self, x_test, y_test, batch_size, num_batches):
        total = num_batches * batch_size
        chunk = max(1, self.EVAL_CHUNK // batch_size) * batch_size
        acc_l = np.zeros(len(self.losses))
        acc_m = [np.zeros(len(row)) for row in self.metrics]
        for lo_i in range(0, total, chunk):
            n = min(chunk, total - lo_i)
            x_mb = [x_test[j][lo_i:lo_i + n] for j in range(len(self.l_in))]
            y_mb = [y_test[j][lo_i:lo_i + n] for j in range(len(self.l_in))]
            y_mb = [DArray.device(v.t) if isinstance(v, _ResidentSlice) else v for v in y_mb]
            self.feed_input(x_mb, key="x_eval")      # own staging buffers: the training step graph keeps its addresses
            self.forward()
            for i in range(len(self.losses)):
                loss = self.losses[i].compute_loss(self.l_out[i].output, y_mb[i])
                self._raise_on_flags(loss, self.losses[i], self.l_out[i].output, y_mb[i], flags=0)
                acc_l[i] += float(loss) * n
            for i, row in enumerate(self.compute_metrics(y_target=y_mb)):
                acc_m[i] += np.asarray(row, dtype=np.float64) * n
        losses = acc_l / float(total)
        metrics = [m / float(total) for m in acc_m]
        return losses, metrics

    # ------------------------------------------------------------------ summary / params / checkpoint
    def summary(self):
        total_params = total_tr = total_notr = 0
        width = 75
        print('=' * width)
        print("Model summary")
        print('=' * width)
        print("{:<20} {:<20} {:<20} {:<10}".format("Layer (type)", "Input Shape/s", "Output Shape", "Param #"))
        print('-' * width)
        for l in self.fts_layers:
            ishapes = [p.oshape for p in l.parents]
            ishapes = ishapes if len(ishapes) > 1 else l.oshape
            tr, notr = l.get_num_params()
            total_tr += tr
            total_notr += notr
            total_params += tr + notr
            print("{:<20} {:<20} {:<20} {:<10}".format(str(l.name), str(ishapes), str(l.oshape), str(tr + notr)))
        print('-' * width)
        print('Total params: {:,}'.format(total_params))
        print('Trainable params: {:,}'.format(total_tr))
        print('Non-trainable params: {:,}'.format(total_notr))
        print('-' * width)
        print('')

    def get_params(self, only_trainable=False):
        """net.py:403-411 — host (float64 ndarray) copies, D2H at the boundary."""
        params = []
        for l in self.fts_layers:
            params.append({k: np.array(np.asarray(v)) for k, v in l.params.items()
                           if (not only_trainable) or (k in l.grads)})
        return params

    def set_params(self, params, only_trainable=False):
        """net.py:413-422 — values are uploaded into the existing (arena) slots."""
        assert len(self.fts_layers) == len(params)
        for i, l in enumerate(self.fts_layers):
            if only_trainable:
                for k in l.params.keys():
                    if k in params[i]:
                        l.params[k] = self._fit_shape(l.params[k], params[i][k])
            else:
                assert len(l.params) == len(params[i])
                for k, v in params[i].items():
                    l.params[k] = self._fit_shape(l.params.get(k), v)

    @staticmethod
    def _fit_shape(cur, value):
        v = np.asarray(value)
        if cur is not None and tuple(v.shape) != cur.shape and v.size == cur.size:
            v = v.reshape(cur.shape)       # e.g. BatchNorm moving stats pickled as (1, *oshape) by the reference
        return v

    def get_grads(self):
        return [{k: np.array(np.asarray(v)) for k, v in l.grads.items()} for l in self.fts_layers]

    def set_grads(self, grads):
        assert len(self.fts_layers) == len(grads)
        for i, l in enumerate(self.fts_layers):
            assert len(l.grads) == len(grads[i])
            for k, v in grads[i].items():
                l.grads[k] = self._fit_shape(l.grads.get(k), v)

    def load(self, filename):
        with open(filename, 'rb') as fp:
            data = pickle.load(fp)
        self.set_params(data['params'])
        grads = data.get('grads')
        if grads:
            self.set_grads(grads)
        print("Model loaded!")

    def save(self, filename, save_grads=False):
        """Same pickle layout as net.py:449-460 (dicts of float64 ndarrays): files are interchangeable
        with the reference's."""
        data = {'params': self.get_params()}
        if save_grads:
            data['grads'] = self.get_grads()
        with open(filename, 'wb') as fp:
            pickle.dump(data, fp)
        print("Model saved!")
