"""Data-parallel parity on the GPU (-m gpu): N ranks shard every minibatch, exchange what the path needs (gradient buckets,
BatchNorm statistics, Embedding mean-delta / last-position, global Dropout masks) and must reproduce the SINGLE-PROCESS
float64 oracle on the global batch — gradients after the first exchange (rel 1e-5) and parameters after three Adam steps
(rel 1e-4) — for all five BASELINE topologies.

N = 2 always; N = 4 and N = 8 when the box has that many GPUs.  With >= N GPUs the ranks use NCCL on separate devices (step
graph captured, collectives inside it); on a 1-GPU box both ranks share cuda:0 and the exchange runs over gloo (same host
logic, same kernels)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [("mnist_mlp", "literal"), ("mnist_cnn", "per_sample"), ("mnist_cnn_fused", "per_sample"), ("iris_mlp", "literal"),
         ("res_conv", "literal"), ("rnn_sentiment", "literal")]


def _world_sizes():
    try:
        import torch
        n = torch.cuda.device_count()
    except Exception:
        n = 0
    return [2] + [w for w in (4, 8) if n >= w]


@pytest.mark.parametrize("world", _world_sizes())
@pytest.mark.parametrize("name,route", CASES, ids=[c[0] for c in CASES])
@pytest.mark.timeout(300)
def test_data_parallel_matches_single_process_oracle(name, route, world):
    import torch
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    env = dict(os.environ, DNB_TEST_BACKEND=backend, MASTER_ADDR="127.0.0.1")
    port = str(29541 + 10 * world + [c[0] for c in CASES].index(name))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", port,
                        os.path.join(ROOT, "tests", "dp_worker.py"), name, route],
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == world
