"""Profiling driver: a few training steps of the MNIST-shape CNN (per_sample routing) at the benchmark batch."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dnpy_b200
from tests import topologies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dnpy_b200.set_math(os.environ.get("MATH", "bf16"))
dnpy_b200.set_maxpool_route(os.environ.get("ROUTE", "per_sample"))
ns = topologies.product_ns()
np.random.seed(42)
net = topologies.mnist_cnn(ns, f1=32, f2=64)
net.set_mode("train")
net.use_graph = False
x, y = topologies.make_data("mnist_cnn", B, np.random.RandomState(1234))
x = torch.from_numpy(x.astype(np.float32)).cuda(); y = torch.from_numpy(y.astype(np.float32)).cuda()
for s in range(steps):
    out = net.train_step([x], [y])
torch.cuda.synchronize()
print("loss", float(out[0][0].item()))
