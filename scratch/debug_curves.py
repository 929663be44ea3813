import sys, numpy as np
sys.path.insert(0, '.')
from tests import topologies
from oracle.oracle_net import OracleNet
import dnpy_b200
ns = topologies.product_ns()

def run(name, kw, batch, steps, per_layer_steps=3):
    print("=====", name, kw, batch)
    np.random.seed(42)
    net = topologies.BUILDERS[name](ns, **kw)
    onet = OracleNet.mirror(net)
    net.set_mode("train"); onet.set_mode("train")
    rs = np.random.RandomState(1234)
    x, y = topologies.make_data(name, batch * 4, rs, **kw)
    st = np.random.get_state()
    np.random.seed(99); s1 = np.random.get_state()
    for s in range(steps):
        lo = (s % 4) * batch
        xb, yb = [x[lo:lo+batch]], [y[lo:lo+batch]]
        np.random.set_state(s1)
        outs = net.train_step(xb, yb)
        got = float(outs[0][0].item())
        np.random.set_state(s1)
        want = onet.step(xb, yb)[0]
        s1 = np.random.get_state()
        line = f"step {s}: loss got {got:.8f} want {want:.8f} rel {abs(got-want)/abs(want):.2e}"
        if s < per_layer_steps:
            print(line)
            for l, n in zip(net.fts_layers, onet.fts):
                errs = []
                if n.output is not None and l.output is not None:
                    a, b = np.asarray(l.output, dtype=float), np.asarray(n.output, dtype=float)
                    errs.append(f"out {np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-6):.1e}")
                if n.delta is not None and l.delta is not None:
                    a, b = np.asarray(l.delta, dtype=float), np.asarray(n.delta, dtype=float)
                    errs.append(f"delta {np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-6):.1e} (max {np.max(np.abs(b)):.1e})")
                for k in n.grads:
                    a, b = np.asarray(l.grads[k], dtype=float).reshape(n.grads[k].shape), n.grads[k]
                    errs.append(f"g_{k} {np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-6):.1e}")
                for k in n.params:
                    a, b = np.asarray(l.params[k], dtype=float).reshape(n.params[k].shape), n.params[k]
                    errs.append(f"p_{k} {np.max(np.abs(a-b))/max(np.max(np.abs(b)),1e-6):.1e}")
                print("   ", l.name, "; ".join(errs))
        elif s % 5 == 0:
            print(line)

run("mnist_cnn", dict(f1=8, f2=16), 16, 12)
run("iris_mlp", dict(), 24, 60, per_layer_steps=1)
