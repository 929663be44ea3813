#!/usr/bin/env python
"""bench.py — CNN train images/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
                    [--math fp32|tf32|bf16] [--batch B] [--route literal|per_sample] [--no-others] [--strong]

A *step* is one pass of the dnpy hot path (net.py:176-195) over one minibatch: feed_input ->
forward -> CrossEntropy loss+delta -> reset_grads -> backward -> [all-reduce] -> Adam.
Workload (config.workload): ``mnist_cnn_wide`` = the MNIST-shape CNN of examples/6_mnist_conv.py
(Conv2D 3x3 none -> MaxPool 3x3/2 -> Relu -> Conv2D 3x3 same -> MaxPool 3x3/2 -> Relu -> Dense 10 ->
Softmax, CE, Adam 0.01) with 32/64 filters, synthetic 1x28x28 inputs, per-GPU batch 8192 (weak
scaling: global batch 8192*N).  ``mnist_cnn_lit`` is the example's literal 2/4 filters.

MaxPool2D routing: ``per_sample`` (default here: the intended scatter, the only routing that shards over GPUs and the one
whose arithmetic is meaningful) — the reference's LITERAL routing (pooling.py:95: every delta into sample 0) is
measured beside it at N=1 (``other_maxpool_route``).
other_workloads: the other BASELINE.json configs (MNIST MLP B=1024, literal 2/4-filter CNN, residual conv net 64x64x3
B=4096, RNN sentiment seq 256) with value / ms_per_step / roofline / cpu_baseline each, at the same N.
strong_scaling (N>1): the same CNN at GLOBAL batch 8192 (per-GPU 8192/N).

value : images/s with the minibatches already resident in HBM (CUDA events, max over ranks).
e2e   : images/s through the public API (Net.train_step) from PINNED HOST buffers, H2D copies of the
        inputs and the D2H read of the loss inside the timed region.
roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md §5.
--impl reference : the reference's own CPU implementation timed on a bounded sample of the same workload, rank 0 only:
        STOCK dnpy (staged by __graft_entry__.build() under the git-ignored baseline/_ref/, imported unmodified with a
        2-function matplotlib stub) when present — kind "reference" — else the vectorised NumPy float64 oracle port
        (kind "port", ~50x faster than the stock per-position Python loops).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# name -> (topology, kwargs, per-GPU batch, batch of the CPU port sample, batch of the stock-reference sample)
WORKLOADS = {
    "mnist_cnn_wide": ("mnist_cnn", dict(f1=32, f2=64), 8192, 64, 8),
    "mnist_cnn_lit": ("mnist_cnn", dict(f1=2, f2=4), 8192, 256, 64),
    "mnist_mlp": ("mnist_mlp", dict(h1=512, h2=256), 1024, 1024, 1024),
    "res_conv": ("res_conv", dict(size=64, c1=32, c2=64), 4096, 4, 0),
    "rnn_sentiment": ("rnn_sentiment", dict(vocab=1000, emb=8, seq=256, hidden=32, trunc=4), 1024, 16, 4),
}
OTHERS = ["mnist_mlp", "mnist_cnn_lit", "res_conv", "rnn_sentiment"]      # the other BASELINE.json configs (cfg2, 3-lit, 4, 5)


def peaks():
    """Measured roofline denominators (driver-written), else the recipe's fallback."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=open(self.path, "w"),
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = dict(sm_mhz=None, sm_max_mhz=None, reasons=[])
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out = dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))
        return out


# ---------------------------------------------------------------------------------------------
# algorithmic work per kernel (DESIGN.md §4 / SURVEY.md §8d): (bytes, flops) for ONE launch
# ---------------------------------------------------------------------------------------------
def algo_work(layer, kernel, B):
    kind = type(layer).__name__
    s = 4
    if kernel in ("conv_pool_act_fwd", "conv_pool_act_fwd_tc", "conv_pool_act_bwd"):
        # fused Conv2D -> MaxPool2D -> Relu block: compulsory traffic = the image, the activation tensor (its delta in
        # backward) and one argmax/gate byte per pooled element; FLOPs = the convolution (forward) / 10 FMAs per pooled
        # element (backward: the routed weight gradient)
        pool, act = layer._cpa
        c, h, w = layer.parents[0].oshape
        f, oh, ow = layer.oshape
        npool = B * int(np.prod(pool.oshape))
        byts = B * c * h * w * s + npool * (s + 1)
        fl = 2.0 * B * f * oh * ow * c * 9 if "fwd" in kernel else 2.0 * npool * 10
        return byts, fl
    if kernel == "tc_conv2d_pool_act_fwd":
        # Conv2D -> MaxPool2D -> LeakyRelu in one tcgen05 launch: the image in, 5 bytes per pooled element out; the FLOPs of
        # the convolution
        pool = layer._tcp
        c, h, w = layer.parents[0].oshape
        f, oh, ow = layer.oshape
        kh, kw = layer.kernel_size[-2:]
        return B * c * h * w * s + B * int(np.prod(pool.oshape)) * (s + 1), 2.0 * B * f * oh * ow * c * kh * kw
    if kind == "Embedding":
        l, d = layer.oshape
        return (B * l * (4 + d * s) if "fwd" in kernel else B * l * (4 + d * s) + layer.input_dim * d * s), 0.0
    if kind == "SimpleRNN":
        t, d = layer.parents[0].oshape
        hd = layer.hidden_dim
        if "fwd" in kernel:
            return B * t * (d + hd) * s, 2.0 * B * t * hd * (hd + d)
        k = min(t, (layer.bptt_truncate if layer.bptt_truncate is not None else t) + 1)
        if kernel == "rnn_bwd_chain":       # the serial chain of the tensor-core backward: states in, pre + dx out
            return B * t * (2 * hd + d) * s, 2.0 * B * t * hd * (hd + d)
        if kernel == "rnn_bwd_lags_tc":     # pre + states + x in; (k - 1) lag products + the reduction against h | x | 1
            return B * t * (2 * hd + d) * s, 2.0 * B * t * hd * ((k - 1) * hd + hd + d + 1)
        # literal truncated BPTT (recurrent.py:228-247): per step the dh chain (2 H*H) + dx (H*D) + k inner steps of
        # gWaa / gWax outer products and the s <- s*(1-h^2) Waa product
        return B * t * (2 * d + hd) * s, 2.0 * B * t * (hd * hd + hd * d + k * (2 * hd * hd + hd * d))
    if kind in ("Conv2D", "PointwiseConv2D"):
        c, h, w = layer.parents[0].oshape
        f, oh, ow = layer.oshape
        kh, kw = layer.kernel_size[-2:]
        nin, nout = B * c * h * w, B * f * oh * ow
        fl = 2.0 * B * f * oh * ow * c * kh * kw
        if "bwd_data" in kernel or "dgrad" in kernel:
            return (nin + nout) * s, fl
        if "bwd_weight" in kernel or "wgrad" in kernel:
            return (nin + nout) * s, fl
        if "reg" in kernel:
            return 3 * f * c * kh * kw * s, 0.0
        if "prep_weights" in kernel:
            return 2 * f * c * kh * kw * s, 0.0
        return (nin + nout) * s, fl
    if kind == "DepthwiseConv2D":
        c, h, w = layer.parents[0].oshape
        _, oh, ow = layer.oshape
        kh, kw = layer.kernel_size[-2:]
        return (B * c * h * w + B * c * oh * ow) * s, 2.0 * B * c * oh * ow * kh * kw
    if kind == "BatchNorm":
        n = B * int(np.prod(layer.oshape))
        return (5 if "bwd" in kernel else 3) * n * s, 0.0          # fwd: 2 reads + 1 write; bwd: x, dy twice, dx
    if kind == "Add":
        n = B * int(np.prod(layer.oshape))
        return (len(layer.parents) + 1) * n * s, 0.0
    if kind == "MaxPool2D":
        c, h, w = layer.parents[0].oshape
        _, oh, ow = layer.oshape
        nin, nout = B * c * h * w, B * c * oh * ow
        if kernel == "maxpool2d_fwd":
            return (nin + 2 * nout) * s, 0.0
        if kernel == "maxpool2d_act_fwd":                   # pool + activation: the planes in, value + byte per pooled element out
            return nin * s + nout * (s + 1), 0.0
        if kernel == "pool_act_bwd":                        # activation delta + byte per pooled element in, the planes out
            return nin * s + nout * (s + 1), 0.0
        if kernel == "maxpool2d_bwd_last_top":
            return min(B, 128) * c * oh * ow * s, 0.0       # argmax offsets of the newest 128 samples
        if kernel == "maxpool2d_bwd_last":
            return nout * s, 0.0                            # upper bound: only unfinished window positions are scanned
        if kernel == "maxpool2d_bwd_literal":
            # LITERAL route: every sample's dx is written (zeros for samples 1..B-1: the memset is part of this entry)
            return (nin + c * h * w + c * oh * ow * 10) * s, 0.0
        return (nin + 2 * nout) * s, 0.0
    if kind in ("Relu", "LeakyRelu"):
        n = B * int(np.prod(layer.oshape))
        return (3 if "bwd" in kernel else 2) * n * s, 0.0
    if kind == "Dense":
        fin, fout = layer.parents[0].oshape[0], layer.units
        fl = 2.0 * B * fin * fout
        if kernel == "dense_bwd_db":
            return (B * fout + 2 * fout) * s, 0.0
        return (B * fin + fin * fout + B * fout) * s, fl
    if kind in ("Softmax", "Sigmoid", "Tanh"):
        n = B * int(np.prod(layer.oshape))
        return (3 if "bwd" in kernel else 2) * n * s, 0.0
    if kind == "Dropout":
        n = B * int(np.prod(layer.oshape))
        return 3 * n * s, 0.0
    return None


def build_net(workload, math, route):
    import dnpy_b200
    from tests import topologies
    dnpy_b200.set_math(math)
    dnpy_b200.set_maxpool_route(route)
    name, kw = WORKLOADS[workload][:2]
    np.random.seed(42)
    ns = topologies.product_ns()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = topologies.BUILDERS[name](ns, **kw)
    net.set_mode("train")
    return net


def make_batches(workload, n_batches, batch, rank):
    from tests import topologies
    name, kw = WORKLOADS[workload][:2]
    rs = np.random.RandomState(1234 + rank)
    xs, ys = [], []
    for _ in range(n_batches):
        x, y = topologies.make_data(name, batch, rs, **kw)
        xs.append(np.ascontiguousarray(x, dtype=np.int32 if x.dtype.kind == "i" else np.float32))
        ys.append(np.ascontiguousarray(y, dtype=np.float32))
    return xs, ys


def _blas_info():
    try:
        from threadpoolctl import threadpool_info
        return [f"{d.get('internal_api')} {d.get('version')} x{d.get('num_threads')}" for d in threadpool_info()]
    except Exception:
        return []


def run_cpu_port(workload, steps, warmup, route, sample_batch=None):
    """The reference's algorithm on the host cores: NumPy float64 oracle port (vectorised over output positions,
    BLAS-threaded — a STRONGER baseline than the reference's own per-position Python loops)."""
    from tests import topologies
    from oracle.oracle_net import OracleNet
    import contextlib, io
    name, kw, _, cpu_batch = WORKLOADS[workload][:4]
    batch = sample_batch or cpu_batch
    np.random.seed(42)
    ns = topologies.product_ns()
    with contextlib.redirect_stdout(io.StringIO()):
        net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net, maxpool_route=route)
    onet.set_mode("train")
    x, y = topologies.make_data(name, batch, np.random.RandomState(1234), **kw)
    for _ in range(max(1, warmup)):
        onet.step([x], [y])
    t0 = time.perf_counter()
    for _ in range(steps):
        onet.step([x], [y])
    dt = (time.perf_counter() - t0) / steps
    return dict(value=batch / dt, ms_per_step=dt * 1e3, batch=batch, cores=os.cpu_count(), blas=_blas_info(), kind="port",
                what="NumPy float64 oracle port (vectorised)")


def stock_reference_ns():
    """The UNMODIFIED reference staged under baseline/_ref (git-ignored; __graft_entry__.build() copies it there in the build
    container, it travels to the GPU box with the snapshot).  dnpy/utils.py:5 imports matplotlib (absent): 2-function stub."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "dnpy")):
        return None
    import types
    if "matplotlib" not in sys.modules:
        mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        plt.imshow = lambda *a, **k: None
        plt.show = lambda *a, **k: None
        mpl.pyplot = plt
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    warnings.filterwarnings("ignore", category=DeprecationWarning)
    import dnpy.layers as RL
    from dnpy import net as rnet, optimizers as ropt, losses as rlosses, metrics as rmetrics, initializers as rinit, regularizers as rreg
    return types.SimpleNamespace(L=RL, Net=rnet.Net, opt=ropt, losses=rlosses, metrics=rmetrics, init=rinit, reg=rreg)


def run_cpu_stock(workload, steps, warmup):
    """One step of dnpy.net.Net.fit's inner loop (net.py:176-195) on STOCK dnpy, through its own public API; the reference's
    MaxPool2D.backward is its literal routing by construction."""
    from tests import topologies
    import contextlib, io
    name, kw, _, _, batch = WORKLOADS[workload]
    rns = stock_reference_ns()
    if rns is None or not batch:
        return None
    np.random.seed(42)
    with contextlib.redirect_stdout(io.StringIO()):
        net = topologies.BUILDERS[name](rns, **kw)
    net.set_mode("train")
    x, y = topologies.make_data(name, batch, np.random.RandomState(1234), **kw)

    def step():
        net.feed_input([x]); net.forward()
        losses, deltas = net.compute_losses(y_target=[y])
        net.reset_grads(); net.do_delta(deltas); net.backward(); net.apply_grads()
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return dict(value=batch / dt, ms_per_step=dt * 1e3, batch=batch, cores=os.cpu_count(), blas=_blas_info(), kind="reference",
                what="stock dnpy (baseline/_ref, unmodified), float64, its own Python loops + NumPy/OpenBLAS")


def cpu_baseline_for(workload, route, port_steps=3, stock_steps=1):
    """cpu_baseline object: the stock reference when it is staged (bounded sample), the oracle port beside it."""
    unit = "samples/s" if WORKLOADS[workload][0] == "rnn_sentiment" else "images/s"
    port = run_cpu_port(workload, port_steps, 1, route)
    stock = None
    try:
        stock = run_cpu_stock(workload, stock_steps, 1)
    except Exception as e:       # the staged copy is optional
        stock = None
    main = stock or port
    out = dict(value=main["value"], unit=unit, cores=main["cores"], kind=main["kind"],
               sample=f"{stock_steps if stock else port_steps} step(s) at batch {main['batch']} ({main['ms_per_step']:.0f} ms/step) of the "
                      f"same topology: {main['what']}; BLAS {main['blas']}")
    if stock:
        out["port"] = dict(value=port["value"], batch=port["batch"], ms_per_step=port["ms_per_step"], what=port["what"])
    return out


class Harness:
    """One workload on this rank's GPU: build, warm up (graph capture), time `value`, time `e2e`, per-kernel roofline."""

    def __init__(self, workload, batch, math, route, rank, world, n_mb=4):
        import torch
        from dnpy_b200 import _lib
        from dnpy_b200.darray import DArray
        self.torch, self.lib, self._lib = torch, _lib.load(), _lib
        self.workload, self.batch, self.math, self.route, self.rank, self.world, self.n_mb = workload, batch, math, route, rank, world, n_mb
        self.net = build_net(workload, math, route)
        xs, ys = make_batches(workload, n_mb, batch, rank)
        is_int = xs[0].dtype.kind == "i"
        self.dx = [DArray.device(torch.from_numpy(x).cuda(), is_int=is_int) for x in xs]
        self.dy = [DArray.device(torch.from_numpy(y).cuda()) for y in ys]
        self.px = [torch.from_numpy(x).pin_memory() for x in xs]
        self.py = [torch.from_numpy(y).pin_memory() for y in ys]
        self.loss_host = [torch.empty((2,), dtype=torch.float32).pin_memory() for _ in range(2)]
        self.loss_evt = [torch.cuda.Event(), torch.cuda.Event()]
        self.last = {}
        self.h2d = int(self.px[0].numel() * self.px[0].element_size() + self.py[0].numel() * self.py[0].element_size())

    def barrier(self):
        if self.world > 1:
            self.torch.distributed.barrier()
        self.torch.cuda.synchronize()

    def timed(self, step_fn, steps):
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = self.lib.dnb_launch_count() + self.net.graph_launches
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        self.barrier()
        ms = e0.elapsed_time(e1)
        launches = self.lib.dnb_launch_count() + self.net.graph_launches - n0
        if self.world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    def step_resident(self, i):
        self.last["o"] = self.net.train_step([self.dx[i % self.n_mb]], [self.dy[i % self.n_mb]])

    def step_e2e(self, i):
        # public API with PINNED HOST minibatches: H2D of x and y, the step, D2H of [loss, flags] — every step.
        # The host reads each loss one step late (the read of step i-1 overlaps step i), so the copy of the
        # next minibatch (dedicated copy stream, double-buffered) overlaps the kernels of the current one.
        outs = self.net.train_step([self.px[i % self.n_mb]], [self.py[i % self.n_mb]])
        self.loss_host[i & 1].copy_(outs[0], non_blocking=True)
        self.loss_evt[i & 1].record()
        if i > 0:
            self.loss_evt[(i - 1) & 1].synchronize()
            self.last["loss"] = float(self.loss_host[(i - 1) & 1][0])

    def warm(self, min_steps, min_seconds=0.5):
        """At least `min_steps` steps, `min_seconds` under load, and enough calls for every (minibatch buffer) step graph
        to be captured (2 eager calls + 1 capture per buffer) so no capture lands in a timed region."""
        torch = self.torch
        t_w, i_w = time.perf_counter(), 0
        need = max(3, min_steps, 3 * self.n_mb + 4)
        while True:
            self.step_resident(i_w)
            i_w += 1
            if i_w % 8 == 0:
                torch.cuda.synchronize()
                more = i_w < need or (time.perf_counter() - t_w) < min_seconds
                if self.world > 1:            # every rank must run the SAME number of steps (collectives inside)
                    flag = torch.tensor([1 if more else 0], device="cuda")
                    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)
                    more = bool(flag.item())
                if not more:
                    return i_w

    def final_loss(self):
        torch = self.torch
        loss = float(self.last["o"][0][0].item())
        if self.world > 1:
            lt = torch.tensor([loss], device="cuda")
            torch.distributed.all_reduce(lt)
            loss = float(lt.item()) / self.world
        if not np.isfinite(loss):
            raise RuntimeError(f"non-finite loss {loss} during the benchmark ({self.workload})")
        return loss

    def measure(self, steps, warmup, e2e=True, min_seconds=0.5):
        i_w = self.warm(warmup, min_seconds)
        ms, launches = self.timed(self.step_resident, steps)
        loss = self.final_loss()
        out = dict(ms=ms, launches=int(launches), loss=loss, warm=i_w, ms_e2e=None)
        if e2e:
            for i in range(10):               # both staging slots: 2 eager calls + capture each
                self.step_e2e(i)
            out["ms_e2e"], _ = self.timed(self.step_e2e, steps)
        return out

    # -------------------------------------------------------------- per-kernel roofline (this rank's view)
    def kernel_table(self, prof_steps):
        """Per-kernel device time from the library's own CUDA-event chain on the launching stream (dnb_prof_begin/end), one
        layer call at a time.  A ~200 us spin kernel first: the CPU enqueues the layer's launches and events while it runs, so
        the chain measures back-to-back GPU time and not the host's launch latency."""
        import ctypes
        from dnpy_b200.darray import stream_ptr
        torch, lib, _lib, net = self.torch, self.lib, self._lib, self.net
        cbuf = ctypes.create_string_buffer(1 << 16)
        breakdown = {}

        def prof_call(tag, layer, fn):
            torch.cuda._sleep(400000)
            _lib.check(lib.dnb_prof_begin(stream_ptr()))
            fn()
            _lib.check(lib.dnb_prof_end(cbuf, len(cbuf)))
            for k, (cnt, tot) in json.loads(cbuf.value.decode()).items():
                e = breakdown.setdefault((tag, k), dict(layer=layer, count=0, ms=0.0))
                e["count"] += cnt
                e["ms"] += tot

        for s in range(prof_steps):
            net.feed_input([self.dx[s % self.n_mb]])
            if hasattr(net, "_prepare_host_draws"):
                pass
            for li, l in enumerate(net.fts_layers):
                prof_call(f"{li}:{l.name}:fwd", l, l.forward)
            outs, deltas = [], []
            for i, lo in enumerate(net.losses):
                prof_call("loss", None, lambda: (outs.append(lo.result_tensor(net.l_out[i].output, self.dy[s % self.n_mb])),
                                                 deltas.append(lo.compute_delta(net.l_out[i].output, self.dy[s % self.n_mb]))))
            prof_call("reset_grads", None, net.reset_grads)
            net.do_delta(deltas)
            for li, l in enumerate(net.bts_layers):
                prof_call(f"{len(net.fts_layers) - 1 - li}:{l.name}:bwd", l, l.backward)
            prof_call("apply_grads", None, net.apply_grads)
        torch.cuda.synchronize()
        return breakdown

    def roofline(self, prof_steps, traffic_tab=None, top_n=24):
        pk = peaks()
        breakdown = self.kernel_table(prof_steps)
        total_ms = sum(e["ms"] for e in breakdown.values()) or 1.0
        rows = sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])
        traffic_tab = traffic_tab or {}
        tensor_peak = pk["bf16_tflops"] if self.math == "bf16" else pk["bf16_tflops"] / 2.0   # TF32 dense = half the bf16 figure
        ridge = tensor_peak * 1e12 / (pk["hbm_gbs"] * 1e9)

        def roofline_of(tag, kname, e):
            per_launch_ms = e["ms"] / max(1, e["count"])
            work = algo_work(e["layer"], kname, self.batch) if e["layer"] is not None else None
            if work is None or per_launch_ms <= 0:
                return None
            byts, flops = work
            # a contraction under the ridge point is accounted against HBM
            if flops > 0 and (flops / byts) > ridge and self.math != "fp32" and ("_sv" in kname or "tc_" in kname):
                ach = flops / (per_launch_ms * 1e-3) / 1e12
                r = dict(bound="tensor", achieved=ach, peak=tensor_peak, unit="TFLOP/s", frac=ach / tensor_peak)
            else:
                ach = byts / (per_launch_ms * 1e-3) / 1e9
                r = dict(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"])
            r.update(traffic=traffic_tab.get(f"{tag}/{kname}", traffic_tab.get(kname)), kernel=f"{tag}/{kname}", ms_per_launch=per_launch_ms,
                     share_of_step=e["ms"] / total_ms, algorithmic_bytes=byts, algorithmic_flops=flops, peak_source=pk["source"])
            if flops > 0:
                r["algorithmic_tflops"] = flops / (per_launch_ms * 1e-3) / 1e12      # fp32 SIMT kernels: of ~72 TFLOP/s nominal
            return r

        (tag, kname), top = rows[0]
        roof = roofline_of(tag, kname, top)
        tops = []
        for (t, k), e in rows[:top_n]:
            r = roofline_of(t, k, e)
            tops.append(dict(kernel=f"{t}/{k}", ms_per_step=e["ms"] / prof_steps, share=e["ms"] / total_ms,
                             bound=r["bound"] if r else None, frac=round(r["frac"], 4) if r else None))
        return roof, tops


def unit_metric(workload):
    name = WORKLOADS[workload][0]
    unit = "samples/s" if name == "rnn_sentiment" else "images/s"
    metric = "CNN train images/sec" if name == "mnist_cnn" else f"{name} train {unit}"
    return name, unit, metric


def describe(workload, batch, world, route, math):
    name, kw = WORKLOADS[workload][:2]
    return dict(workload=f"{workload}{kw}", topology=name, per_gpu_batch=batch, global_batch=batch * world,
                optimizer="Adam", loss="CrossEntropy" if name != "rnn_sentiment" else "BinaryCrossEntropy",
                parallelism=f"dp{world}", maxpool_route=route,
                l2_policy="4 distinct minibatches cycled; per-step activations exceed the 126 MB L2 on the CNN / residual configs "
                          "(the MLP / RNN working sets are L2-resident by nature: 3-25 MB per step)",
                math={"fp32": "fp32 SIMT everywhere", "tf32": "TF32 operands / fp32 accumulate on the tcgen05 conv + dense kernels, fp32 elsewhere",
                      "bf16": "BF16 operands / fp32 accumulate on the tcgen05 shifted-view conv kernels (TF32 on other tensor-core "
                              "kernels), fp32 storage and fp32 arithmetic everywhere else"}[math])


def main():
    # stdout carries exactly ONE JSON line: anything a library prints (NCCL banner ...) goes to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--workload", default="mnist_cnn_wide", choices=list(WORKLOADS))
    ap.add_argument("--math", default=os.environ.get("DNPY_B200_BENCH_MATH", "bf16"), choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--route", default="per_sample", choices=["literal", "per_sample"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the informational run with the other MaxPool routing")
    ap.add_argument("--no-others", action="store_true", help="skip the other BASELINE configs (other_workloads)")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling figure (N>1)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    def_batch = WORKLOADS[args.workload][2]
    batch = args.batch or def_batch
    name, unit, metric = unit_metric(args.workload)
    route = args.route
    config = describe(args.workload, batch, world, route, args.math)

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        r = None
        try:
            r = run_cpu_stock(args.workload, steps, warm)
        except Exception:
            r = None
        if r is None:
            r = run_cpu_port(args.workload, steps, max(1, warm), "literal")
        line = dict(impl="reference", metric=metric, value=r["value"], unit=unit, n_gpus=args.gpus, steps=steps,
                    warmup=warm, ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(config, per_gpu_batch=r["batch"], global_batch=r["batch"], maxpool_route="literal",
                                math="float64 NumPy (the reference's only dtype)"),
                    cpu_baseline=dict(value=r["value"], unit=unit, cores=r["cores"], kind=r["kind"],
                                      sample=f"{steps} step(s) at batch {r['batch']} of the same topology: {r['what']}; BLAS {r['blas']}"),
                    e2e=dict(value=r["value"], unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
        return 0

    # ---------------------------------------------------------------- native arm (GPU)
    import torch
    from dnpy_b200 import parallel
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the backend has no CPU fallback)")
    if world > 1:
        parallel.init()
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.cuda.current_device()

    sampler = ClockSampler(dev)
    sampler.start()                      # nvidia-smi needs ~100 ms to start: begin before the warm-up
    h = Harness(args.workload, batch, args.math, route, rank, world)
    m = h.measure(args.steps, args.warmup)
    clocks = sampler.stop()
    ms, ms_e2e = m["ms"], m["ms_e2e"]
    value = batch * world * args.steps / (ms * 1e-3)
    e2e_value = batch * world * args.steps / (ms_e2e * 1e-3)

    roof, tops = None, None
    if rank == 0 or world > 1:           # layers with collectives inside (sync-BatchNorm, Embedding) need every rank to take part
        traffic_tab = {}
        tpath = os.path.join(ROOT, "profiles", "traffic_r02.json")      # dram bytes per launch from the ncu --set full capture
        if os.path.exists(tpath) and args.workload == "mnist_cnn_wide" and batch == def_batch:
            traffic_tab = json.load(open(tpath)).get(f"{args.math}/{route}", {})
        roof, tops = h.roofline(max(2, min(5, args.steps)), traffic_tab)
    if world > 1:
        torch.distributed.barrier()
    # ---------------------------------------------------------------- the other MaxPool routing (N=1 only, informational)
    # LITERAL is dnpy/layers/pooling.py:95 as written (every sample's delta lands in sample 0); PER_SAMPLE is the intended
    # scatter-add, the variant whose arithmetic is meaningful and the only one that shards.  Same protocol.
    alt = None
    if world == 1 and name == "mnist_cnn" and not args.no_alt:
        other = "per_sample" if route == "literal" else "literal"
        ha = Harness(args.workload, batch, args.math, other, rank, world)
        ma = ha.measure(args.steps, args.warmup, e2e=False, min_seconds=0.2)
        alt = dict(maxpool_route=other, value=batch * args.steps / (ma["ms"] * 1e-3), unit=unit, ms_per_step=ma["ms"] / args.steps)
        del ha
    # ---------------------------------------------------------------- strong scaling (N>1): global batch fixed at 8192
    strong = None
    if world > 1 and name == "mnist_cnn" and not args.no_strong and def_batch % world == 0:
        hs = Harness(args.workload, def_batch // world, args.math, route, rank, world)
        msr = hs.measure(args.steps, args.warmup, e2e=False, min_seconds=0.2)
        strong = dict(scaling="strong", global_batch=def_batch, per_gpu_batch=def_batch // world,
                      value=def_batch * args.steps / (msr["ms"] * 1e-3), unit=unit, ms_per_step=msr["ms"] / args.steps)
        del hs
    # ---------------------------------------------------------------- the other BASELINE configs, same N
    others = None
    if not args.no_others and args.workload == "mnist_cnn_wide" and not args.batch:
        others = {}
        for wl in OTHERS:
            wname, wunit, wmetric = unit_metric(wl)
            wb = WORKLOADS[wl][2]
            torch.cuda.empty_cache()
            ho = Harness(wl, wb, args.math, route, rank, world, n_mb=2)
            mo = ho.measure(min(args.steps, 20), args.warmup, e2e=True, min_seconds=0.2)
            ent = dict(metric=wmetric, unit=wunit, value=wb * world * min(args.steps, 20) / (mo["ms"] * 1e-3),
                       ms_per_step=mo["ms"] / min(args.steps, 20),
                       e2e=dict(value=wb * world * min(args.steps, 20) / (mo["ms_e2e"] * 1e-3), unit=wunit,
                                h2d_bytes_per_step=ho.h2d, d2h_bytes_per_step=8),
                       gpu_launches=mo["launches"], final_loss=mo["loss"], config=describe(wl, wb, world, route, args.math))
            if rank == 0 or world > 1:
                r_o, t_o = ho.roofline(2, None, top_n=8)
                ent["roofline"], ent["top_kernels"] = r_o, t_o
            if rank == 0:
                if world == 1 and not args.no_cpu:
                    ent["cpu_baseline"] = cpu_baseline_for(wl, route, port_steps=2, stock_steps=1)
            if world > 1:
                torch.distributed.barrier()
            others[wl] = ent
            del ho
    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_for(args.workload, route)

    if rank == 0:
        line = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=args.math, data="synthetic",
                    config=dict(config, warmup_steps_run=m["warm"], cuda_graph=bool(h.net.use_graph and h.net._graph_ok_for_world())),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=unit, h2d_bytes_per_step=h.h2d, d2h_bytes_per_step=8,
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=m["launches"], roofline=roof, cpu_baseline=cpu, top_kernels=tops, final_loss=m["loss"],
                    other_maxpool_route=alt, strong_scaling=strong, other_workloads=others)
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        # every rank has finished its collectives; tear down without the NCCL destructor dance (captured step graphs
        # still reference the communicator, and a blocking destroy_process_group has been seen to wait forever)
        torch.distributed.barrier()
        torch.cuda.synchronize()
        real_stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())
