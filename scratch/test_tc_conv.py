import sys, numpy as np, torch
sys.path.insert(0, '.')
import dnpy_b200
from tests import topologies
ns = topologies.product_ns()
L = ns.L
def run(cin, h, w, f, k, s, pad, B, kind="Conv2D"):
    res = {}
    rs = np.random.RandomState(0)
    x = rs.randn(B, cin, h, w); 
    for math in ("fp32", "tf32"):
        dnpy_b200.set_math(math)
        l0 = L.Input(shape=(cin, h, w)); l0.output = x
        if kind == "Conv2D":
            l1 = L.Conv2D(l0, f, kernel_size=k, strides=s, padding=pad)
        else:
            l1 = L.PointwiseConv2D(l0, f)
        np.random.seed(3); l1.initialize()
        l1.params['w1'] = np.random.RandomState(1).randn(*l1.params['w1'].shape) * 0.3
        l1.params['b1'] = np.random.RandomState(2).randn(*l1.params['b1'].shape) * 0.3
        l1.forward()
        out = np.asarray(l1.output).copy()
        d = np.random.RandomState(4).randn(*out.shape)
        for g in l1.grads.values(): g.fill(0.0)
        l1.delta = d
        l1.backward()
        torch.cuda.synchronize()
        res[math] = (out, np.asarray(l0.delta).copy(), np.asarray(l1.grads['w1']).copy(), np.asarray(l1.grads['b1']).copy())
    names = ("out", "dx", "gw", "gb")
    msg = []
    ok = True
    for n, a, b in zip(names, res["tf32"], res["fp32"]):
        e = np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-6)
        ok &= e < 5e-3
        msg.append(f"{n} {e:.1e}")
    print(f"{kind} C={cin} {h}x{w} F={f} k={k} s={s} {pad} B={B}: " + "; ".join(msg) + ("  OK" if ok else "  FAIL"), flush=True)
run(32, 12, 12, 64, (3,3), (1,1), "same", 16)
run(1, 28, 28, 32, (3,3), (1,1), "none", 8)
run(8, 12, 12, 16, (3,3), (1,1), "same", 5)
run(3, 32, 32, 32, (3,3), (1,1), "same", 4)
run(32, 16, 16, 64, (3,3), (2,2), "same", 4)
run(32, 16, 16, 32, None, None, None, 4, kind="PointwiseConv2D")
run(2, 8, 8, 3, (2,2), (2,2), "none", 3)
run(16, 10, 12, 130, (3,3), (1,1), "same", 3)
run(64, 8, 8, 32, (3,3), (1,1), "same", 9)
print("done")
