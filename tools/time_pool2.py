"""Times the MaxPool2D(3x3/2) -> Relu pair kernels of the CNN's second block (64x12x12 planes) through the C ABI."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dnpy_b200 import _lib
from dnpy_b200._lib import Win
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
lib = _lib.load()
gp = Win(B, 64, 12, 12, 64, 5, 5, 3, 3, 2, 2, 0, 0, 0, 0)
dev = "cuda"
y = torch.randn(B, 64, 12, 12, device=dev); act = torch.empty(B, 64, 5, 5, device=dev)
ig = torch.empty(B * 64 * 25, dtype=torch.uint8, device=dev)
po = torch.empty(B, 64, 5, 5, device=dev); pi = torch.empty(B, 64, 5, 5, dtype=torch.int32, device=dev)
dy = torch.empty_like(y); dact = torch.randn(B, 64, 5, 5, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
p = lambda t: C.c_void_p(t.data_ptr())
def t(fn):
    fn(); torch.cuda.synchronize()
    tot = 0.0
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / n * 1e3
print("maxpool2d_fwd us", round(t(lambda: _lib.call("dnb_maxpool2d_fwd", p(y), p(po), p(pi), C.byref(gp), st)), 1))
print("maxpool2d_act_fwd us", round(t(lambda: _lib.call("dnb_maxpool2d_act_fwd", p(y), p(act), p(ig), C.byref(gp), 0.0, st)), 1))
print("pool_act_bwd us", round(t(lambda: _lib.call("dnb_pool_act_bwd", p(dact), p(ig), p(dy), C.byref(gp), 0.0, st)), 1))
print("maxpool2d_bwd us", round(t(lambda: _lib.call("dnb_maxpool2d_bwd", p(dact), p(pi), p(dy), C.byref(gp), 1, None, st)), 1))
big = torch.empty(B, 64, 12, 12, device=dev)
print("memset 302MB us", round(t(lambda: big.zero_()), 1), " fill_ us", round(t(lambda: big.fill_(1.5)), 1))
print("read-only sum 302MB us", round(t(lambda: big.sum()), 1))
big2 = torch.empty_like(big)
print("copy 302MB->302MB us", round(t(lambda: big2.copy_(big)), 1))
