"""Times conv2 of the CNN (32x12x12 -> 64, 3x3 same) through the C ABI alone: plain forward, the per-role clock profile
(DNB_SV_PROF), the debug switches (DNB_SV_DEBUG), the fused pool forward, dgrad/wgrad and dnb_pool_act_bwd."""
import sys, os, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dnpy_b200 import _lib
from dnpy_b200._lib import Win, Reg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
what = sys.argv[2] if len(sys.argv) > 2 else "all"
lib = _lib.load()
g = Win(B, 32, 12, 12, 64, 12, 12, 3, 3, 1, 1, 1, 1, 1, 1)
gp = Win(B, 64, 12, 12, 64, 5, 5, 3, 3, 2, 2, 0, 0, 0, 0)
dev = "cuda"
x = torch.randn(B, 32, 12, 12, device=dev); w = torch.randn(64, 32, 3, 3, device=dev) * 0.05; b = torch.randn(64, device=dev)
y = torch.empty(B, 64, 12, 12, device=dev); act = torch.empty(B, 64, 5, 5, device=dev); ig = torch.empty(B * 64 * 25, dtype=torch.uint8, device=dev)
dy = torch.randn(B, 64, 12, 12, device=dev); dx = torch.empty_like(x); gw = torch.zeros_like(w); gb = torch.zeros_like(b)
dact = torch.randn(B, 64, 5, 5, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ws = torch.empty(max(1, lib.dnb_conv2d_ws_bytes(C.byref(g), 2)), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: C.c_void_p(t.data_ptr())
def t(fn, n=10):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
fwd = lambda: _lib.call("dnb_conv2d_fwd", p(x), p(w), p(b), p(y), C.byref(g), p(ws), 2, st)
fpool = lambda: _lib.call("dnb_conv2d_pool_act_fwd", p(x), p(w), p(b), p(act), p(ig), C.byref(g), C.byref(gp), 0.0, p(ws), 2, st)
bwd = lambda: _lib.call("dnb_conv2d_bwd", p(x), p(w), p(b), p(dy), p(dx), p(gw), p(gb), C.byref(g), 1.0 / B, Reg(0, 0), Reg(0, 0), 1.0, p(ws), 2, st)
pab = lambda: _lib.call("dnb_pool_act_bwd", p(dact), p(ig), p(dy), C.byref(gp), 0.0, st)
print("fwd us", round(t(fwd), 1))
print("fwd+pool us", round(t(fpool), 1))
print("bwd (dgrad+wgrad) us", round(t(bwd), 1))
print("pool_act_bwd us", round(t(pab), 1))
for dbg in (1, 2, 8, 3, 9, 10, 11):
    os.environ["DNB_SV_DEBUG"] = str(dbg)
    print("fwd debug", dbg, "us", round(t(fwd), 1))
del os.environ["DNB_SV_DEBUG"]
os.environ["DNB_SV_PROF"] = "1"
fwd(); fpool()
torch.cuda.synchronize()
