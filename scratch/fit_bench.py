"""Wall-clock of Net.fit itself (the user-facing loop) on a resident synthetic set: images/s including every host-side cost."""
import sys, time, io, contextlib
sys.path.insert(0, '.')
import numpy as np, torch
import bench
net = bench.build_net("mnist_cnn_wide", "bf16", "literal")
from tests import topologies
rs = np.random.RandomState(0)
N, B = 8192 * 8, 8192
x, y = topologies.make_data("mnist_cnn", N, rs, f1=32, f2=64)
x = x.astype(np.float32); y = y.astype(np.float32)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    net.fit([x], [y], batch_size=B, epochs=3, evaluate_epoch=False, print_rate=100)      # warm-up: upload, eager steps, graph capture
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    E = 10
    net.fit([x], [y], batch_size=B, epochs=E, evaluate_epoch=False, print_rate=100)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"Net.fit: {E} epochs x {N} samples in {dt*1e3:.1f} ms -> {E*N/dt/1e6:.2f} M images/s ({dt/E/(N//B)*1e3:.3f} ms/step incl. the per-fit dataset upload)")
with contextlib.redirect_stdout(buf):
    t0 = time.perf_counter()
    net.fit([x], [y], batch_size=B, epochs=1, evaluate_epoch=True, print_rate=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
print(f"one printed epoch + evaluate(batch_size=1) over {N} samples: {dt*1e3:.1f} ms")
