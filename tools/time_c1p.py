"""Times dnb_conv_pool_act_fwd_tc alone (CUDA events, 20 launches) at the benchmark batch; argv: B F."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dnpy_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
F = int(sys.argv[2]) if len(sys.argv) > 2 else 32
lib = _lib.load()
x = torch.rand(B, 1, 28, 28, device="cuda"); w = torch.randn(F, 1, 3, 3, device="cuda") * 0.4; b = torch.randn(F, device="cuda") * 0.1
act = torch.empty(B, F, 12, 12, device="cuda"); ig = torch.empty(B, F, 12, 12, dtype=torch.uint8, device="cuda")
g = _lib.Win(B, 1, 28, 28, F, 26, 26, 3, 3, 1, 1, 0, 0, 0, 0); gp = _lib.Win(B, F, 26, 26, F, 12, 12, 3, 3, 2, 2, 0, 0, 0, 0)
def run(): _lib.check(lib.dnb_conv_pool_act_fwd_tc(x.data_ptr(), w.data_ptr(), b.data_ptr(), act.data_ptr(), ig.data_ptr(), g, gp, 0.0, _lib.MATH_BF16, None))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("c1p_fwd_tc us", round(e0.elapsed_time(e1) / 20 * 1e3, 1), "B", B, "F", F)
