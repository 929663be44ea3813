"""Data-parallel parity on the GPU (-m gpu): two ranks shard every minibatch, all-reduce the flat
gradient arena and must reproduce the SINGLE-PROCESS float64 oracle on the global batch — gradients
after the first exchange (rel 1e-5) and parameters after three Adam steps (rel 1e-4).

With >= 2 GPUs the ranks use NCCL on separate devices; on a 1-GPU box both ranks share cuda:0 and the
exchange runs over gloo (same host logic, same kernels)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name,route", [("mnist_mlp", "literal"), ("mnist_cnn", "per_sample")])
@pytest.mark.timeout(300)
def test_two_rank_data_parallel_matches_single_process_oracle(name, route):
    import torch
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    env = dict(os.environ, DNB_TEST_BACKEND=backend, MASTER_ADDR="127.0.0.1")
    port = "29541" if name == "mnist_mlp" else "29542"
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", port,
                        os.path.join(ROOT, "tests", "dp_worker.py"), name, route],
                       capture_output=True, text=True, env=env, timeout=280)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2
