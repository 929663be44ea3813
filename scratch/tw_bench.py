import os, sys, json, ctypes, numpy as np, torch
sys.path.insert(0, '.')
import dnpy_b200
from dnpy_b200 import _lib
from dnpy_b200.darray import DArray, stream_ptr
from tests import topologies
ns = topologies.product_ns(); L = ns.L
lib = _lib.load()
dnpy_b200.set_math("bf16")
B = 8192
l0 = L.Input(shape=(1, 28, 28)); l0.output = DArray.device(torch.randn(B, 1, 28, 28, device="cuda"))
l1 = L.Conv2D(l0, 32, kernel_size=(3, 3), strides=(1, 1), padding="none"); np.random.seed(3); l1.initialize(); l1.forward()
l1.delta = DArray.device(torch.randn(B, *l1.oshape, device="cuda"))
cbuf = ctypes.create_string_buffer(1 << 16)
def prof(fn, n=5):
    fn(); torch.cuda.synchronize(); agg = {}
    for _ in range(n):
        torch.cuda._sleep(400000)
        _lib.check(lib.dnb_prof_begin(stream_ptr())); fn(); _lib.check(lib.dnb_prof_end(cbuf, len(cbuf)))
        for k, (cnt, tot) in json.loads(cbuf.value.decode()).items(): agg[k] = agg.get(k, 0.0) + tot / n
    return {k: round(v * 1e3, 1) for k, v in agg.items()}
for dbg in sys.argv[1:] or ["0"]:
    os.environ["DNB_TW_DEBUG"] = dbg
    print("debug", dbg, prof(l1.backward), flush=True)
