import sys, time, numpy as np, torch
sys.path.insert(0, '.')
import bench, dnpy_b200
from dnpy_b200.darray import DArray
from dnpy_b200.layers.base import Layer
net = bench.build_net("mnist_cnn_wide", "tf32", "literal")
xs, ys = bench.make_batches("mnist_cnn_wide", 4, 8192, 0)
dx = [DArray.device(torch.from_numpy(x).cuda()) for x in xs]
dy = [DArray.device(torch.from_numpy(y).cuda()) for y in ys]
for i in range(30):
    torch.cuda.synchronize(); t = time.perf_counter()
    net.train_step([dx[i % 4]], [dy[i % 4]])
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(i, f"{dt*1e3:.2f} ms gen {Layer.buffer_generation} graphs {len(net._graphs)} captured {sum(1 for e in net._graphs.values() if e['graph'] is not None)}")
