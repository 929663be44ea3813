"""One EAGER training step of a bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`).
    python scratch/prof_step.py [workload] [math] [route]"""
import os, sys
os.environ["DNPY_B200_GRAPH"] = "0"
sys.path.insert(0, ".")
import numpy as np, torch
import bench
wl = sys.argv[1] if len(sys.argv) > 1 else "mnist_cnn_wide"
math = sys.argv[2] if len(sys.argv) > 2 else "tf32"
route = sys.argv[3] if len(sys.argv) > 3 else "literal"
torch.cuda.set_device(0)
from dnpy_b200 import _lib
from dnpy_b200.darray import DArray
lib = _lib.load()
net = bench.build_net(wl, math, route)
batch = int(os.environ.get("PROF_BATCH", bench.WORKLOADS[wl][2]))
xs, ys = bench.make_batches(wl, 2, batch, 0)
is_int = xs[0].dtype.kind == "i"
dx = [DArray.device(torch.from_numpy(x).cuda(), is_int=is_int) for x in xs]
dy = [DArray.device(torch.from_numpy(y).cuda()) for y in ys]
for i in range(3):
    net.train_step([dx[i % 2]], [dy[i % 2]])
torch.cuda.synchronize()
n0 = lib.dnb_launch_count()
torch.cuda.profiler.start()
out = net.train_step([dx[1]], [dy[1]])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("launches/step", lib.dnb_launch_count() - n0, "loss", float(out[0][0].item()))
