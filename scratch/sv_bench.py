"""Time the conv2 layer of mnist_cnn_wide (C=32 12x12 -> F=64, 3x3 same, B=8192) per kernel, with the
shifted-view kernel's debug switches (1 = no epilogue stores, 2 = no MMAs, 4 = no loads)."""
import os, sys, json, ctypes, numpy as np, torch
sys.path.insert(0, '.')
import dnpy_b200
from dnpy_b200 import _lib
from dnpy_b200.darray import DArray, stream_ptr
from tests import topologies
ns = topologies.product_ns(); L = ns.L
lib = _lib.load()
dnpy_b200.set_math(os.environ.get("SVMATH", "tf32"))
B = int(os.environ.get("SVB", "8192"))
def setup(cin, h, w, f, k, s, pad):
    l0 = L.Input(shape=(cin, h, w))
    l0.output = DArray.device(torch.randn(B, cin, h, w, device="cuda"))
    l1 = L.Conv2D(l0, f, kernel_size=k, strides=s, padding=pad)
    np.random.seed(3); l1.initialize()
    l1.forward()
    l1.delta = DArray.device(torch.randn(B, *l1.oshape, device="cuda"))
    return l0, l1
cbuf = ctypes.create_string_buffer(1 << 16)
def prof(fn, n=5):
    fn(); torch.cuda.synchronize()
    agg = {}
    for _ in range(n):
        _lib.check(lib.dnb_prof_begin(stream_ptr())); fn(); _lib.check(lib.dnb_prof_end(cbuf, len(cbuf)))
        for k, (cnt, tot) in json.loads(cbuf.value.decode()).items():
            agg[k] = agg.get(k, 0.0) + tot / n
    return {k: round(v * 1e3, 1) for k, v in agg.items()}
l0, l1 = setup(32, 12, 12, 64, (3, 3), (1, 1), "same")
for dbg in sys.argv[1:] or ["0"]:
    os.environ["DNB_SV_DEBUG"] = dbg
    print("debug", dbg, "fwd", prof(l1.forward), "bwd", prof(l1.backward), flush=True)
