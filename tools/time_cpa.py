"""Times the fused block kernels alone (CUDA events, 20 launches each) at the benchmark batch."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dnpy_b200
from dnpy_b200 import _lib
from tests import topologies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dnpy_b200.set_math("bf16"); dnpy_b200.set_maxpool_route("per_sample")
ns = topologies.product_ns()
np.random.seed(42)
net = topologies.mnist_cnn(ns, f1=32, f2=64); net.set_mode("train"); net.use_graph = False
x, y = topologies.make_data("mnist_cnn", B, np.random.RandomState(1234))
x = torch.from_numpy(x.astype(np.float32)).cuda(); y = torch.from_numpy(y.astype(np.float32)).cuda()
for _ in range(2): net.train_step([x], [y])
conv = net.fts_layers[1]
def t(fn, n=20):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n * 1e3
net.feed_input([x])
print("cpa_fwd us", round(t(conv.forward), 1))
net.forward(); lo = net.losses[0]; d = lo.compute_delta(net.l_out[0].output, dnpy_b200.darray.DArray.device(y))
net.reset_grads(); net.do_delta([d])
for l in net.bts_layers[:-2]: l.backward()
def bwd():
    conv._cpa_live = live; conv.backward()
conv.forward(); live = dict(conv._cpa_live)
print("cpa_bwd us", round(t(bwd), 1))
