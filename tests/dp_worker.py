"""Worker of tests/test_gpu_dp.py: one data-parallel rank (launched by torchrun)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import dnpy_b200                                   # noqa: E402
from dnpy_b200 import parallel                     # noqa: E402
from oracle.oracle_net import OracleNet            # noqa: E402
from tests import topologies                       # noqa: E402


def main():
    name, route = sys.argv[1], sys.argv[2]
    rank, world = parallel.init(backend=os.environ.get("DNB_TEST_BACKEND", "gloo"))
    dnpy_b200.set_math("fp32")
    dnpy_b200.set_maxpool_route(route)
    kw = dict(mnist_mlp=dict(h1=32, h2=16), mnist_cnn=dict(f1=4, f2=8))[name]
    ns = topologies.product_ns()
    np.random.seed(42)                              # identical replicas on every rank
    net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net, maxpool_route=route)
    net.set_mode("train"); onet.set_mode("train")
    batch, steps = 16, 3
    x, y = topologies.make_data(name, batch * steps, np.random.RandomState(7), **kw)
    for s in range(steps):
        x_mb, y_mb = net.get_minibatch([x], [y], s, batch)        # this rank's rows of the global minibatch
        assert len(x_mb[0]) == batch // world
        net.train_step(x_mb, y_mb)
        onet.step([x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]])    # single process, global batch
        if s == 0:
            for l, n in zip(net.fts_layers, onet.fts):
                for k, g in n.grads.items():
                    a = np.asarray(l.grads[k]).reshape(g.shape)
                    err = np.max(np.abs(a - g)) / max(np.max(np.abs(g)), 1e-6)
                    assert err <= 1e-5, (rank, l.name, k, err)
    for l, n in zip(net.fts_layers, onet.fts):
        for k, v in n.params.items():
            a = np.asarray(l.params[k]).reshape(v.shape)
            err = np.max(np.abs(a - v)) / max(np.max(np.abs(v)), 1e-3)
            assert err <= 1e-4, (rank, l.name, k, err)
    parallel.shutdown()
    print(f"dp rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
