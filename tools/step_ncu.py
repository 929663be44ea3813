"""One eager training step of the benchmark CNN between cudaProfilerStart/Stop (ncu --profile-from-start off):
   ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/step_r02 python tools/step_ncu.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dnpy_b200
from tests import topologies
math = sys.argv[1] if len(sys.argv) > 1 else "bf16"
route = sys.argv[2] if len(sys.argv) > 2 else "per_sample"
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8192
dnpy_b200.set_math(math); dnpy_b200.set_maxpool_route(route)
ns = topologies.product_ns()
np.random.seed(42)
net = topologies.mnist_cnn(ns, f1=32, f2=64); net.set_mode("train"); net.use_graph = False
x, y = topologies.make_data("mnist_cnn", B, np.random.RandomState(1234))
x = torch.from_numpy(x.astype(np.float32)).cuda(); y = torch.from_numpy(y.astype(np.float32)).cuda()
for _ in range(3): net.train_step([x], [y])
torch.cuda.synchronize()
torch.cuda.profiler.start()
net.train_step([x], [y])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
