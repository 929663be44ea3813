#!/usr/bin/env python
"""bench.py — CNN train images/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]
                    [--math fp32|tf32] [--batch B] [--route literal|per_sample]

A *step* is one pass of the dnpy hot path (net.py:176-195) over one minibatch: feed_input ->
forward -> CrossEntropy loss+delta -> reset_grads -> backward -> [all-reduce] -> Adam.
Workload (config.workload): ``mnist_cnn_wide`` = the MNIST-shape CNN of examples/6_mnist_conv.py
(Conv2D 3x3 none -> MaxPool 3x3/2 -> Relu -> Conv2D 3x3 same -> MaxPool 3x3/2 -> Relu -> Dense 10 ->
Softmax, CE, Adam 0.01) with 32/64 filters, synthetic 1x28x28 inputs, per-GPU batch 8192 (weak
scaling: global batch 8192*N).  ``mnist_cnn_lit`` is the example's literal 2/4 filters.

value : images/s with the minibatches already resident in HBM (CUDA events, max over ranks).
e2e   : images/s through the public API (Net.train_step) from PINNED HOST buffers, H2D copies of the
        inputs and the D2H read of the loss inside the timed region.
roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md §5.
--impl reference : the reference's CPU algorithm (oracle port, NumPy float64 + OpenBLAS, all host
        threads) timed on a bounded sample of the same workload; rank 0 only.
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

WORKLOADS = {
    "mnist_cnn_wide": ("mnist_cnn", dict(f1=32, f2=64), 8192, 64),
    "mnist_cnn_lit": ("mnist_cnn", dict(f1=2, f2=4), 8192, 256),
    "mnist_mlp": ("mnist_mlp", dict(h1=512, h2=256), 1024, 1024),
    "res_conv": ("res_conv", dict(size=64, c1=32, c2=64), 512, 4),
    "rnn_sentiment": ("rnn_sentiment", dict(vocab=1000, emb=8, seq=256, hidden=32, trunc=4), 1024, 16),
}


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
    if kind == "Softmax":
        n = B * int(np.prod(layer.oshape))
        return 2 * n * s, 0.0
    return None


def build_net(workload, math, route):
    import dnpy_b200
    from tests import topologies
    dnpy_b200.set_math(math)
    dnpy_b200.set_maxpool_route(route)
    name, kw, _, _ = WORKLOADS[workload]
    np.random.seed(42)
    ns = topologies.product_ns()
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        net = topologies.BUILDERS[name](ns, **kw)
    net.set_mode("train")
    return net


def make_batches(workload, n_batches, batch, rank):
    from tests import topologies
    name, kw, _, _ = WORKLOADS[workload]
    rs = np.random.RandomState(1234 + rank)
    xs, ys = [], []
    for _ in range(n_batches):
        x, y = topologies.make_data(name, batch, rs, **kw)
        xs.append(np.ascontiguousarray(x, dtype=np.int32 if x.dtype.kind == "i" else np.float32))
        ys.append(np.ascontiguousarray(y, dtype=np.float32))
    return xs, ys


def run_cpu_reference(workload, steps, warmup, sample_batch=None):
    """The reference's algorithm on the host cores: NumPy float64 oracle port (vectorised over output
    positions, BLAS-threaded — a STRONGER baseline than the reference's own per-position Python
    loops, whose cost model `oracle.*_loops` keeps)."""
    from tests import topologies
    from oracle.oracle_net import OracleNet
    import contextlib, io
    name, kw, _, cpu_batch = WORKLOADS[workload]
    batch = sample_batch or cpu_batch
    np.random.seed(42)
    ns = topologies.product_ns()
    with contextlib.redirect_stdout(io.StringIO()):
        net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net)
    onet.set_mode("train")
    x, y = topologies.make_data(name, batch, np.random.RandomState(1234), **kw)
    for _ in range(max(1, warmup)):
        onet.step([x], [y])
    t0 = time.perf_counter()
    for _ in range(steps):
        onet.step([x], [y])
    dt = (time.perf_counter() - t0) / steps
    try:
        from threadpoolctl import threadpool_info
        blas = [f"{d.get('internal_api')} {d.get('version')} x{d.get('num_threads')}" for d in threadpool_info()]
    except Exception:
        blas = []
    return dict(value=batch / dt, ms_per_step=dt * 1e3, batch=batch, cores=os.cpu_count(), blas=blas)


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
    ap.add_argument("--route", default="literal", choices=["literal", "per_sample"])
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the workload's)")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-alt", action="store_true", help="skip the informational run with the other MaxPool routing")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    name, kw, def_batch, cpu_batch = WORKLOADS[args.workload]
    batch = args.batch or def_batch
    unit = "samples/s" if name == "rnn_sentiment" else "images/s"
    metric = "CNN train images/sec" if name == "mnist_cnn" else f"{name} train {unit}"
    config = dict(workload=f"{args.workload}{kw}", topology=name, per_gpu_batch=batch, global_batch=batch * world,
                  optimizer="Adam", loss="CrossEntropy" if name != "rnn_sentiment" else "BinaryCrossEntropy",
                  parallelism=f"dp{world}", maxpool_route=args.route,
                  l2_policy="per-step activations (>1 GB) exceed the 126 MB L2; 4 distinct minibatches cycled",
                  math={"fp32": "fp32 SIMT everywhere", "tf32": "TF32 operands / fp32 accumulate on the tcgen05 conv + dense kernels, fp32 elsewhere",
                        "bf16": "BF16 operands / fp32 accumulate on the tcgen05 shifted-view conv kernels (TF32 on other tensor-core "
                                "kernels), fp32 storage and fp32 arithmetic everywhere else"}[args.math])

    # ---------------------------------------------------------------- reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return 0
        steps = max(1, min(args.steps, 5))
        r = run_cpu_reference(args.workload, steps, min(args.warmup, 1))
        line = dict(impl="reference", metric=metric, value=r["value"], unit=unit, n_gpus=args.gpus, steps=steps,
                    warmup=min(args.warmup, 1), ms_per_step=r["ms_per_step"], higher_is_better=True, scaling="weak",
                    vs_baseline=None, dtype="f64", data="synthetic", config=dict(config, per_gpu_batch=r["batch"],
                                                                                 global_batch=r["batch"]),
                    cpu_baseline=dict(value=r["value"], unit=unit, cores=r["cores"], kind="port",
                                      sample=f"{steps} steps at batch {r['batch']} of the same topology, NumPy float64 "
                                             f"oracle port, BLAS {r['blas']}"),
                    e2e=dict(value=r["value"], unit=unit, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                    gpu_launches=0)
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
        return 0

    # ---------------------------------------------------------------- native arm (GPU)
    import torch
    from dnpy_b200 import _lib, parallel
    from dnpy_b200.darray import DArray, stream_ptr
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the backend has no CPU fallback)")
    if world > 1:
        parallel.init()
    else:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dev = torch.cuda.current_device()
    lib = _lib.load()
    net = build_net(args.workload, args.math, args.route)
    n_mb = 4
    xs, ys = make_batches(args.workload, n_mb, batch, rank)
    is_int = xs[0].dtype.kind == "i"
    # device-resident minibatches for `value`
    dx = [DArray.device(torch.from_numpy(x).cuda(), is_int=is_int) for x in xs]
    dy = [DArray.device(torch.from_numpy(y).cuda()) for y in ys]
    # pinned host minibatches for `e2e`
    px = [torch.from_numpy(x).pin_memory() for x in xs]
    py = [torch.from_numpy(y).pin_memory() for y in ys]
    loss_host = [torch.empty((2,), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_evt = [torch.cuda.Event(), torch.cuda.Event()]

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = lib.dnb_launch_count() + net.graph_launches
        e0.record()
        for i in range(steps):
            step_fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        launches = lib.dnb_launch_count() + net.graph_launches - n0
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    last = {}

    def step_resident(i):
        last["o"] = net.train_step([dx[i % n_mb]], [dy[i % n_mb]])

    def step_e2e(i):
        # public API with PINNED HOST minibatches: H2D of x and y, the step, D2H of [loss, flags] — every step.
        # The host reads each loss one step late (the read of step i-1 overlaps step i), so the copy of the
        # next minibatch (dedicated copy stream, double-buffered) overlaps the kernels of the current one.
        outs = net.train_step([px[i % n_mb]], [py[i % n_mb]])
        loss_host[i & 1].copy_(outs[0], non_blocking=True)
        loss_evt[i & 1].record()
        if i > 0:
            loss_evt[(i - 1) & 1].synchronize()
            last["loss"] = float(loss_host[(i - 1) & 1][0])

    sampler = ClockSampler(dev)
    sampler.start()                      # nvidia-smi needs ~100 ms to start: begin before the warm-up
    t_w = time.perf_counter()
    i_w = 0
    # warm-up: at least W steps, at least 0.5 s under load, and enough calls for every (minibatch buffer) step
    # graph to be captured (2 eager calls + 1 capture per buffer) so no capture lands in the timed region
    min_warm = max(3, args.warmup, 3 * n_mb + 4)
    while True:
        step_resident(i_w)
        i_w += 1
        if i_w % 8 == 0:
            torch.cuda.synchronize()
            more = i_w < min_warm or (time.perf_counter() - t_w) < 0.5
            if world > 1:                # every rank must run the SAME number of steps (one all-reduce each)
                flag = torch.tensor([1 if more else 0], device="cuda")
                torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MAX)
                more = bool(flag.item())
            if not more:
                break
    ms, launches = timed(step_resident, args.steps)
    loss = float(last["o"][0][0].item())
    if world > 1:
        lt = torch.tensor([loss], device="cuda")
        torch.distributed.all_reduce(lt)
        loss = float(lt.item()) / world
    if not np.isfinite(loss):
        raise RuntimeError(f"non-finite loss {loss} during the benchmark")
    for i in range(10):                  # both staging slots: 2 eager calls + capture each
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop()
    value = batch * world * args.steps / (ms * 1e-3)
    e2e_value = batch * world * args.steps / (ms_e2e * 1e-3)
    h2d = int(px[0].numel() * px[0].element_size() + py[0].numel() * py[0].element_size())

    # ---------------------------------------------------------------- per-kernel roofline (rank 0 view)
    roof = None
    breakdown = {}
    if rank == 0:
        pk = peaks()
        buf = (b"\0" * (1 << 16))
        import ctypes
        cbuf = ctypes.create_string_buffer(1 << 16)
        prof_steps = max(2, min(5, args.steps))

        def prof_call(tag, layer, fn):
            # a ~200 us spin kernel first: the CPU enqueues the layer's launches and events while it runs, so the event
            # chain measures back-to-back GPU time and not the host's launch latency (which tripled the apparent cost of
            # every kernel under ~30 us)
            torch.cuda._sleep(400000)
            _lib.check(lib.dnb_prof_begin(stream_ptr()))
            fn()
            _lib.check(lib.dnb_prof_end(cbuf, len(cbuf)))
            for k, (cnt, tot) in json.loads(cbuf.value.decode()).items():
                e = breakdown.setdefault((tag, k), dict(layer=layer, count=0, ms=0.0))
                e["count"] += cnt
                e["ms"] += tot

        for s in range(prof_steps):
            net.feed_input([dx[s % n_mb]])
            for li, l in enumerate(net.fts_layers):
                prof_call(f"{li}:{l.name}:fwd", l, l.forward)
            outs, deltas = [], []
            for i, lo in enumerate(net.losses):
                prof_call("loss", None, lambda: (outs.append(lo.result_tensor(net.l_out[i].output, dy[s % n_mb])),
                                                 deltas.append(lo.compute_delta(net.l_out[i].output, dy[s % n_mb]))))
            prof_call("reset_grads", None, net.reset_grads)
            net.do_delta(deltas)
            for li, l in enumerate(net.bts_layers):
                prof_call(f"{len(net.fts_layers) - 1 - li}:{l.name}:bwd", l, l.backward)
            prof_call("apply_grads", None, net.apply_grads)
        torch.cuda.synchronize()
        total_ms = sum(e["ms"] for e in breakdown.values()) or 1.0
        rows = sorted(breakdown.items(), key=lambda kv: -kv[1]["ms"])
        traffic_tab = {}
        tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")      # dram bytes per launch from the ncu --set full capture
        if os.path.exists(tpath) and args.workload == "mnist_cnn_wide" and batch == def_batch:
            traffic_tab = json.load(open(tpath)).get(args.math, {})
        tensor_peak = pk["bf16_tflops"] if args.math == "bf16" else pk["bf16_tflops"] / 2.0   # TF32 dense = half the bf16 figure
        ridge = tensor_peak * 1e12 / (pk["hbm_gbs"] * 1e9)

        def roofline_of(tag, kname, e):
            per_launch_ms = e["ms"] / max(1, e["count"])
            work = algo_work(e["layer"], kname, batch) if e["layer"] is not None else None
            if work is None or per_launch_ms <= 0:
                return None
            byts, flops = work
            # a contraction under the ridge point is accounted against HBM
            if flops > 0 and (flops / byts) > ridge and args.math != "fp32" and "_sv" in kname:
                ach = flops / (per_launch_ms * 1e-3) / 1e12
                r = dict(bound="tensor", achieved=ach, peak=tensor_peak, unit="TFLOP/s", frac=ach / tensor_peak)
            else:
                ach = byts / (per_launch_ms * 1e-3) / 1e9
                r = dict(bound="hbm", achieved=ach, peak=pk["hbm_gbs"], unit="GB/s", frac=ach / pk["hbm_gbs"])
            r.update(traffic=traffic_tab.get(f"{tag}/{kname}"), kernel=f"{tag}/{kname}", ms_per_launch=per_launch_ms,
                     share_of_step=e["ms"] / total_ms, algorithmic_bytes=byts, algorithmic_flops=flops, peak_source=pk["source"])
            return r

        (tag, kname), top = rows[0]
        roof = roofline_of(tag, kname, top)
        top5 = []
        for (t, k), e in rows[:24]:
            r = roofline_of(t, k, e)
            top5.append(dict(kernel=f"{t}/{k}", ms_per_step=e["ms"] / prof_steps, share=e["ms"] / total_ms,
                             bound=r["bound"] if r else None, frac=round(r["frac"], 4) if r else None))
    # ---------------------------------------------------------------- the other MaxPool routing (N=1 only, informational)
    # LITERAL (default) is dnpy/layers/pooling.py:95 as written (every sample's delta lands in sample 0); PER_SAMPLE is the
    # intended scatter-add, the variant whose arithmetic is meaningful.  Same protocol, K timed steps.
    alt = None
    if world == 1 and name == "mnist_cnn" and not args.no_alt:
        import dnpy_b200
        other = "per_sample" if args.route == "literal" else "literal"
        net_main = net
        try:
            net = build_net(args.workload, args.math, other)
            for i in range(3 * n_mb + 8):
                step_resident(i)
            torch.cuda.synchronize()
            ms_alt, _ = timed(step_resident, args.steps)
            alt = dict(maxpool_route=other, value=batch * args.steps / (ms_alt * 1e-3), unit=unit, ms_per_step=ms_alt / args.steps)
        finally:
            net = net_main
            dnpy_b200.set_maxpool_route(args.route)
    # ---------------------------------------------------------------- CPU baseline (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        r = run_cpu_reference(args.workload, args.cpu_steps, 1)
        cpu = dict(value=r["value"], unit=unit, cores=r["cores"], kind="port",
                   sample=f"{args.cpu_steps} steps at batch {r['batch']} ({r['ms_per_step']:.0f} ms/step) of the same "
                          f"topology with the NumPy float64 oracle port (vectorised; BLAS {r['blas']})")

    if rank == 0:
        line = dict(metric=metric, value=value, unit=unit, n_gpus=world, steps=args.steps, warmup=max(3, args.warmup),
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None,
                    dtype=args.math, data="synthetic", config=dict(config, warmup_steps_run=i_w, cuda_graph=bool(net.use_graph and net._graph_ok_for_world())),
                    clocks=clocks,
                    e2e=dict(value=e2e_value, unit=unit, h2d_bytes_per_step=h2d, d2h_bytes_per_step=8,
                             ms_per_step=ms_e2e / args.steps),
                    gpu_launches=int(launches), roofline=roof, cpu_baseline=cpu, top_kernels=top5, final_loss=loss,
                    other_maxpool_route=alt)
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
