"""Data parallelism: one process per GPU (torchrun), minibatch rows sharded over the ranks, ONE
all-reduce per step on the flat gradient arena (NCCL over NVLink 5 / NVSwitch).

The reference has no distributed code; this is the natural sharding of ``Net.fit``'s minibatch
(SURVEY.md §8e).  Scaling rules that keep the all-reduced SUM equal to the single-process value:
  * Dense / Conv / PRelu gradients are batch MEANS in the reference -> kernels use 1/(B_local*world);
  * BatchNorm and SimpleRNN gradients are batch SUMS -> plain sum;
  * regulariser terms are batch independent -> each rank adds 1/world of them;
  * the loss scalar is the mean of per-rank means (equal shards).
Not exchanged (documented in DESIGN.md): BatchNorm batch statistics, Embedding's last-write-wins
scatter and MaxPool's literal routing are per-rank under data parallelism.
"""
import os

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
    """The single exchange step of the data-parallel path: between backward and apply_grads."""
    a = net._bind_arenas()
    allreduce_flat(a["G"])
    for l, k in a["slots"]:
        l.grads[k].written()
