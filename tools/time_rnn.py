"""Runs dnb_rnn_fwd alone at the benchmark shape (B=1024, T=256, D=8, H=32); argv: B T."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dnpy_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
D, H = 8, 32
lib = _lib.load()
x = torch.randn(B, T, D, device="cuda"); waa = torch.randn(H, H, device="cuda") * 0.1; wax = torch.randn(H, D, device="cuda") * 0.1
ba = torch.randn(H, device="cuda") * 0.1; st = torch.empty(B, T, H, device="cuda"); y = torch.empty(B, H, device="cuda")
def run(): _lib.check(lib.dnb_rnn_fwd(x.data_ptr(), waa.data_ptr(), wax.data_ptr(), ba.data_ptr(), st.data_ptr(), y.data_ptr(), B, T, D, H, None))
for _ in range(3): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): run()
e1.record(); torch.cuda.synchronize()
print("rnn_fwd us", round(e0.elapsed_time(e1) / 20 * 1e3, 1), "B", B, "T", T)
