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

SMALL = dict(mnist_mlp=dict(h1=32, h2=16), mnist_cnn=dict(f1=4, f2=8), mnist_cnn_fused=dict(f1=32, f2=64), iris_mlp=dict(),
             res_conv=dict(size=8, c1=4, c2=8), rnn_sentiment=dict(vocab=50, emb=3, seq=7, hidden=5, trunc=2))


def main():
    name, route = sys.argv[1], sys.argv[2]
    rank, world = parallel.init(backend=os.environ.get("DNB_TEST_BACKEND", "gloo"))
    dnpy_b200.set_math("fp32")
    dnpy_b200.set_maxpool_route(route)
    kw = SMALL[name]
    topo = "mnist_cnn" if name == "mnist_cnn_fused" else name
    ns = topologies.product_ns()
    np.random.seed(42)                              # identical replicas on every rank
    net = topologies.BUILDERS[topo](ns, **kw)
    # the SINGLE-PROCESS float64 oracle on the global minibatch (per-sample pool routing: the literal routing does not shard)
    onet = OracleNet.mirror(net, maxpool_route="per_sample")
    net.set_mode("train"); onet.set_mode("train")
    batch, steps = (24 if name == "iris_mlp" else 16), 3
    x, y = topologies.make_data(topo, batch * steps, np.random.RandomState(7), **kw)
    np.random.seed(1234)                            # the shared host stream Dropout draws from (same state on every rank)
    # 32/64 filters at Adam lr 0.01 is chaotic within two steps (the loss jumps to ~10): that case follows the oracle's
    # trajectory (parameters re-synchronised every step, gradients compared at EVERY step) instead of running free
    teacher = name == "mnist_cnn_fused"
    for s in range(steps):
        if teacher:
            net.set_params([dict(n.params) for n in onet.fts])
        x_mb, y_mb = net.get_minibatch([x], [y], s, batch)        # this rank's rows of the global minibatch
        assert len(x_mb[0]) == batch // world
        state = np.random.get_state()
        net.train_step(x_mb, y_mb)
        after = np.random.get_state()
        np.random.set_state(state)                  # the oracle draws the SAME global masks
        onet.step([x[s * batch:(s + 1) * batch]], [y[s * batch:(s + 1) * batch]])    # single process, global batch
        np.random.set_state(after)
        if s == 0 or teacher:
            gmax = max([float(np.max(np.abs(g))) for n in onet.fts for g in n.grads.values()] + [1e-30])
            for l, n in zip(net.fts_layers, onet.fts):
                for k, g in n.grads.items():
                    a = np.asarray(l.grads[k]).reshape(g.shape)
                    err = np.max(np.abs(a - g)) / max(np.max(np.abs(g)), 1e-3 * gmax)
                    # later teacher-forced steps: an fp32-vs-float64 pooling near-tie moves ~1e-5 of a filter's gradient
                    assert err <= (1e-5 if s == 0 else 1e-4), (rank, s, l.name, k, err)
    for l, n in zip(net.fts_layers, onet.fts):
        for k, v in n.params.items():
            if teacher:
                break
            a = np.asarray(l.params[k]).reshape(v.shape)
            if k in n.grads and float(np.max(np.abs(n.grads[k]))) < 1e-12:
                continue                            # a gradient that is pure round-off (bias in front of a BatchNorm)
            err = np.max(np.abs(a - v)) / max(np.max(np.abs(v)), 1e-3)
            assert err <= 1e-4, (rank, l.name, k, err)
    parallel.shutdown()
    print(f"dp rank {rank}/{world} ok")


if __name__ == "__main__":
    main()
