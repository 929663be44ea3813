"""Debug helper: fused conv->pool->act forward/backward vs the layer-by-layer kernels on the GPU."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dnpy_b200
from tests import topologies
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ns = topologies.product_ns()
outs = {}
for fused in (True, False):
    if not fused:
        os.environ["DNPY_B200_NO_CPA"] = "1"
    dnpy_b200.set_maxpool_route("per_sample")
    np.random.seed(42)
    net = topologies.mnist_cnn(ns, f1=32, f2=64)
    net.set_mode("train")
    for l in net.fts_layers:
        if "b1" in l.params:
            l.params["b1"] = np.random.RandomState(5).randn(*l.params["b1"].shape) * 0.1
    x, y = topologies.make_data("mnist_cnn", B, np.random.RandomState(1234))
    net.feed_input([x]); net.forward()
    lo, de = net.compute_losses([y])
    net.reset_grads(); net.do_delta(de); net.backward()
    outs[fused] = dict(act=np.asarray(net.fts_layers[3].output).copy(), idx=np.asarray(net.fts_layers[2].output_idxs).copy(),
                       gw=np.asarray(net.fts_layers[1].grads["w1"]).copy(), gb=np.asarray(net.fts_layers[1].grads["b1"]).copy(), loss=lo[0])
a, b = outs[True], outs[False]
print("loss", a["loss"], b["loss"])
d = np.abs(a["act"] - b["act"])
print("act maxdiff", d.max(), "frac bad", np.mean(d > 1e-5))
bad = np.argwhere(d > 1e-5)
print("first bad", bad[:10].tolist())
for ax, nm in enumerate("bfyx"):
    print(nm, "bad counts per index", np.bincount(bad[:, ax], minlength=a["act"].shape[ax]).tolist()[:40])
print("idx mismatch", np.mean(a["idx"] != b["idx"]))
print("gw relerr", np.abs(a["gw"] - b["gw"]).max() / np.abs(b["gw"]).max(), "gb relerr", np.abs(a["gb"] - b["gb"]).max() / np.abs(b["gb"]).max())
