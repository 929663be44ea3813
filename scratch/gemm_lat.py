import sys, torch
sys.path.insert(0, '.')
from dnpy_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
def t(M, N, K, a_mn=0, b_mn=0, reps=20):
    A = torch.randn(K, M, device="cuda") if a_mn else torch.randn(M, K, device="cuda")
    B = torch.randn(K, N, device="cuda") if b_mn else torch.randn(N, K, device="cuda")
    C = torch.empty(M, N, device="cuda")
    for _ in range(3): _lib.check(lib.dnb_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, a_mn, b_mn, None))
    torch.cuda.synchronize()
    torch.cuda._sleep(2000000)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): _lib.check(lib.dnb_gemm_tf32(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, a_mn, b_mn, None))
    e1.record(); torch.cuda.synchronize()
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: {e0.elapsed_time(e1)/reps*1e3:.1f} us")
for K in (32, 128, 512, 784, 2048):
    t(1024, 512, K)
t(1024, 512, 784, 0, 1)
t(784, 512, 1024, 1, 1)
t(1024, 784, 512, 0, 0)
t(128, 64, 784)
t(8192, 512, 784)
