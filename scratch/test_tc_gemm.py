import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from dnpy_b200 import _lib
lib = _lib.load()
torch.manual_seed(0)
def run(M, N, K, a_mn, b_mn):
    A = torch.randn(M, K, device='cuda'); B = torch.randn(K, N, device='cuda')
    a_store = A.t().contiguous() if a_mn else A.contiguous()          # [K,M] or [M,K]
    b_store = B.contiguous() if b_mn else B.t().contiguous()          # [K,N] or [N,K]
    Cc = torch.full((M, N), float('nan'), device='cuda')
    rc = lib.dnb_gemm_tf32(a_store.data_ptr(), b_store.data_ptr(), Cc.data_ptr(), M, N, K, a_mn, b_mn, None)
    if rc != 0:
        print(M, N, K, a_mn, b_mn, "rc", rc, lib.dnb_last_error()); return
    torch.cuda.synchronize()
    ref = (A.double() @ B.double())
    err = (Cc.double() - ref).abs().max().item() / ref.abs().max().item()
    print(f"M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.2e} {'OK' if err < 5e-3 else 'FAIL'}", flush=True)
for a_mn in (0, 1):
    for b_mn in (0, 1):
        run(128, 128, 32, a_mn, b_mn)
        run(128, 64, 64, a_mn, b_mn)
        run(256, 32, 96, a_mn, b_mn)
        run(1024, 512, 784, a_mn, b_mn)
        run(200, 100, 52, a_mn, b_mn)
        run(1000, 36, 40, a_mn, b_mn)
print("done")
