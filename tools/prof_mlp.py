"""MLP step (cfg2, B=1024) with the GEMM stage profile (DNB_GEMM_PROF) and per-kernel event times."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dnpy_b200
from tests import topologies
math = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dnpy_b200.set_math(math); dnpy_b200.set_maxpool_route("per_sample")
ns = topologies.product_ns()
np.random.seed(42)
net = topologies.mnist_mlp(ns); net.set_mode("train")
x, y = topologies.make_data("mnist_mlp", 1024, np.random.RandomState(1))
x = torch.from_numpy(x.astype(np.float32)).cuda(); y = torch.from_numpy(y.astype(np.float32)).cuda()
for _ in range(4): net.train_step([x], [y])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): net.train_step([x], [y])
e1.record(); torch.cuda.synchronize()
print("graph step us", e0.elapsed_time(e1) / 50 * 1e3)
net.use_graph = False
os.environ["DNB_GEMM_PROF"] = "1"
net.train_step([x], [y])
torch.cuda.synchronize()
