"""Data parallelism: one process per GPU (torchrun), minibatch rows sharded over the ranks, ONE
all-reduce per step on the flat gradient arena (NCCL over NVLink 5 / NVSwitch).

The reference has no distributed code; this is the natural sharding of ``Net.fit``'s minibatch
(SURVEY.md §8e).  Scaling rules that keep the all-reduced SUM equal to the single-process value:
  * Dense / Conv / PRelu gradients are batch MEANS in the reference -> kernels use 1/(B_local*world);
  * BatchNorm and SimpleRNN gradients are batch SUMS -> plain sum;
  * regulariser terms are batch independent -> each rank adds 1/world of them;
  * the loss scalar is the mean of per-rank means (equal shards).
Exchanged besides the gradients, so that N ranks compute exactly what one process computes on the global
minibatch: BatchNorm's batch statistics (forward: every rank's (mean, M2); backward: the two
column sums), Embedding's batch-mean delta (SUM) and last-position table (MAX); Dropout / GaussianNoise draw
the global block from the shared NumPy stream and keep their rows.  MaxPool2D's LITERAL routing (every delta
lands in global sample 0, pooling.py:95) does not shard: with world > 1 the PER_SAMPLE routing is used.
The gradient all-reduce is issued in buckets from inside backward (reverse layer order = arena order), so
all but the last bucket overlap the remaining backward kernels.
"""
import os

from . import _lib
from .runtime import config


def init(backend=None):
    """Join the torchrun process group (reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if torch.cuda.is_available():
            # one GPU per rank; ranks wrap around when a test box has fewer GPUs than ranks (gloo only)
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")) % torch.cuda.device_count())
        dist.init_process_group(backend=backend)
    config.world = dist.get_world_size()
    config.rank = dist.get_rank()
    if config.world > 1 and config.maxpool_route == _lib.ROUTE_LITERAL:
        # pooling.py:95 routes EVERY sample's delta into sample 0 of the (global) minibatch: that does not shard
        if config.rank == 0:
            print("[INFO] data parallel: MaxPool2D uses the per-sample routing (the literal routing does not shard)")
    return config.rank, config.world


def shutdown():
    import torch.distributed as dist
    if dist.is_initialized():
        dist.destroy_process_group()
    config.world, config.rank = 1, 0


def shard_bounds(n, rank=None, world=None):
    """Contiguous equal row blocks: rank r owns [r*n/world, (r+1)*n/world) (remainder dropped,
    like Net.fit drops the last partial minibatch, net.py:162)."""
    rank = config.rank if rank is None else rank
    world = config.world if world is None else world
    per = n // world
    return rank * per, (rank + 1) * per


def shard_rows(a):
    lo, hi = shard_bounds(len(a))
    return a[lo:hi]


def allreduce_flat(tensor):
    """Sum-all-reduce of a flat buffer in place (any backend)."""
    import torch.distributed as dist
    if config.world > 1:
        dist.all_reduce(tensor, op=dist.ReduceOp.SUM)
    return tensor


def allreduce(tensor, op="sum", async_op=False):
    """In-place all-reduce of a device tensor ("sum" or "max"); returns the work handle when async."""
    import torch.distributed as dist
    if config.world == 1:
        return None
    return dist.all_reduce(tensor, op=dist.ReduceOp.SUM if op == "sum" else dist.ReduceOp.MAX, async_op=async_op)


def gather_slots(slots):
    """All-gather through a SUM: ``slots`` is a zeroed [world, ...] device tensor in which this rank filled row
    ``config.rank``; adding the ranks' tensors leaves every row in place on every rank (x + 0 is exact), on any
    backend and inside a captured step graph."""
    return allreduce(slots, "sum")


def mean_over_ranks(values):
    """Average a short list of Python floats over the ranks (equal shards => global mean)."""
    import torch
    import torch.distributed as dist
    if config.world == 1:
        return values
    t = torch.tensor([float(v) for v in values], dtype=torch.float64,
                     device="cuda" if (torch.cuda.is_available() and dist.get_backend() == "nccl") else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(v) / config.world for v in t.tolist()]


def allreduce_grads(net):
    """The whole gradient arena in one all-reduce (the un-bucketed form; Net.backward issues buckets instead)."""
    a = net._bind_arenas()
    allreduce_flat(a["G"])
    for l, k in a["slots"]:
        l.grads[k].written()


class _RawCuda:
    """A raw device allocation behind the CUDA array interface (torch.as_tensor wraps it without a copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = dict(shape=(int(n),), typestr="<f4", data=(int(ptr), False), version=2)


class PeerArena:
    """The flat gradient arena in CUDA-IPC memory that every rank of the node maps, and the one-kernel exchange over it
    (csrc/dp_peer.cu: flags + 16-byte loads through NVLink / NVSwitch peer memory, no NCCL call on the step path).
    ``tensor`` is this rank's arena as a torch tensor; ``allreduce()`` leaves the global sum in it on every rank."""

    def __init__(self, nfloats):
        import ctypes as C
        import torch
        import torch.distributed as dist
        lib = _lib.load()
        self.n = (int(nfloats) + 3) // 4 * 4
        self.rank, self.world = config.rank, config.world

        def alloc(nbytes):
            ptr, handle = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(lib.dnb_ipc_alloc(C.byref(ptr), nbytes, handle))
            return ptr.value, handle.raw
        self._own = [alloc(self.n * 4), alloc(int(lib.dnb_dp_flag_bytes()))]
        gathered = [None] * self.world
        dist.all_gather_object(gathered, [h for _, h in self._own])
        self._opened = []
        ptrs = [[], []]
        for r in range(self.world):
            for j in range(2):
                if r == self.rank:
                    ptrs[j].append(self._own[j][0])
                else:
                    p = C.c_void_p()
                    _lib.check(lib.dnb_ipc_open(gathered[r][j], C.byref(p)))
                    self._opened.append(p.value)
                    ptrs[j].append(p.value)
        self._bufs = (C.c_void_p * self.world)(*ptrs[0])
        self._flags = (C.c_void_p * self.world)(*ptrs[1])
        self._raw = _RawCuda(self._own[0][0], self.n)
        self.tensor = torch.as_tensor(self._raw, device="cuda")
        self.scratch = torch.empty(self.n, dtype=torch.float32, device="cuda")
        dist.barrier()                              # every rank has mapped every arena before the first exchange

    def close(self):
        """Unmap the peers' arenas and free this rank's (every rank must call it at the same point)."""
        import torch
        import torch.distributed as dist
        lib = _lib.load()
        torch.cuda.synchronize()
        dist.barrier()                              # nobody is still reading this arena
        for p in self._opened:
            lib.dnb_ipc_close(p)
        self._opened = []
        dist.barrier()
        for p, _ in self._own:
            lib.dnb_ipc_free(p)
        self._own = []

    def allreduce(self):
        from .darray import stream_ptr
        _lib.call("dnb_dp_allreduce", self._bufs, self._flags, self.scratch.data_ptr(), self.n, self.rank, self.world,
                  stream_ptr())


def peer_exchange_possible():
    """All ranks on one node, NCCL backend (one GPU per rank or a shared test GPU), not switched off (DNPY_B200_DP_PEER=0)."""
    import socket
    import torch
    import torch.distributed as dist
    if config.world < 2 or config.world > 8 or os.environ.get("DNPY_B200_DP_PEER", "1") == "0" or not torch.cuda.is_available():
        return False
    hosts = [None] * config.world
    dist.all_gather_object(hosts, socket.gethostname())
    return len(set(hosts)) == 1


class GradBuckets:
    """Bucketed, overlapped gradient exchange.  The arena is laid out in backward order (Net._trainable walks
    bts_layers), so the gradients of the layers that have ALREADY run backward form a growing prefix: after a layer
    that owns parameters, the finished, not yet sent range [sent, end) goes out as one asynchronous all-reduce
    (it runs on the process group's communication stream, after the kernels queued so far, concurrently with the
    backward kernels that follow).  ``finish`` sends the remainder and makes the compute stream wait for all of
    them before the optimizer.  Buckets smaller than ``min_bytes`` are merged into the next one, except the last."""

    def __init__(self, net, min_bytes=16 << 10):
        a = net._bind_arenas()
        self.G = a["G"]
        self.end_of = {}
        pos = 0
        for (l, k), nxt in zip(a["slots"], a["offsets"][1:] + [a["total"]]):
            self.end_of[id(l)] = nxt          # the arena position just past this layer's last slot
        self.total = a["total"]
        self.min_floats = max(1, min_bytes // 4)
        self.sent = 0
        self.works = []

    def layer_done(self, layer):
        end = self.end_of.get(id(layer))
        if end is None or end - self.sent < self.min_floats:
            return
        self.works.append(allreduce(self.G[self.sent:end], "sum", async_op=True))
        self.sent = end

    def finish(self):
        if self.sent < self.total:
            self.works.append(allreduce(self.G[self.sent:self.total], "sum", async_op=True))
            self.sent = self.total
        for w in self.works:
            if w is not None:
                w.wait()
        self.works = []
