"""CPU tests of the host-side logic: DArray semantics, ParamDict rebinding, topological order,
shape/padding arithmetic (incl. the reference's quirks), sharding maths and a 2-rank gloo run of
the flat-gradient exchange."""
import os
import pickle
import subprocess
import sys
import textwrap

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_darray_behaves_like_ndarray_on_host():
    from dnpy_b200.darray import DArray, ParamDict
    d = DArray(host=np.arange(6).reshape(2, 3))
    assert d.shape == (2, 3) and d.size == 6 and d.ndim == 2
    assert np.all(d == np.arange(6).reshape(2, 3))
    assert np.all(np.arange(6).reshape(2, 3) == d)
    d.fill(0.0)
    assert float(np.abs(np.asarray(d)).sum()) == 0.0
    d[0, 1] = 5
    assert np.asarray(d)[0, 1] == 5.0
    assert np.asarray(d).dtype == np.float64
    back = pickle.loads(pickle.dumps(d))
    assert isinstance(back, np.ndarray) and back[0, 1] == 5.0
    pd = ParamDict(w=np.zeros((2, 2)))
    pd["w"] = np.ones((2, 2))
    assert isinstance(pd["w"], DArray) and float(np.asarray(pd["w"]).sum()) == 4.0


def test_topological_order_and_multi_consumer(ns):
    L = ns.L
    l_in = L.Input(shape=(4,))
    l1 = L.Relu(L.Dense(l_in, 4))
    l2 = L.Relu(L.Dense(l1, 4))
    l = L.Add([l1, l2])
    l_out = L.Softmax(L.Dense(l, 3))
    net = ns.Net()
    net.build([l_in], [l_out], ns.opt.Adam(), [ns.losses.CrossEntropy()], [[ns.metrics.CategoricalAccuracy()]])
    names = [x.name for x in net.fts_layers]
    assert names[0] == "Input" and names[-1] == "Softmax"
    pos = {id(x): i for i, x in enumerate(net.fts_layers)}
    for x in net.fts_layers:
        for p in x.parents:
            assert pos[id(p)] < pos[id(x)]
    assert net.bts_layers == net.fts_layers[::-1]
    # every layer owns an independent optimizer copy with state for EVERY param (Q14)
    opts = {id(x.optimizer) for x in net.fts_layers}
    assert len(opts) == len(net.fts_layers)


def test_cycle_raises(ns):
    L = ns.L
    a = L.Input(shape=(3,))
    b = L.Relu(a)
    a.parents.append(b)
    net = ns.Net()
    with pytest.raises(RecursionError):
        net.build([a], [b], ns.opt.SGD(), [ns.losses.MSE()], [[ns.metrics.MSE()]])


def test_shape_quirks(ns):
    L = ns.L
    i = L.Input(shape=(3, 9, 9))
    c = L.Conv2D(i, 2, kernel_size=(3, 3), strides=(2, 2), padding="same")
    assert c.oshape == (2, 5, 5) and c.pads == (2, 2) and (c.pad_top, c.pad_bottom) == (1, 1)
    c2 = L.Conv2D(L.Input(shape=(1, 8, 8)), 2, kernel_size=(2, 2), strides=(1, 1), padding="same")
    assert (c2.pad_top, c2.pad_bottom, c2.pad_left, c2.pad_right) == (0, 1, 0, 1)     # more pad bottom/right
    # Q4: pooling derives both pads from the WIDTH of the output
    p = L.MaxPool2D(L.Input(shape=(1, 6, 8)), pool_size=(3, 3), strides=(1, 1), padding="same")
    assert tuple(int(v) for v in p.pads) == (4, 2)       # probed on the reference: height pad from the width
    with pytest.raises(ValueError):
        L.Dense(L.Input(shape=(2, 2)), 3)
    with pytest.raises(ValueError):
        L.Add([L.Input(shape=(2,)), L.Input(shape=(3,))])
    with pytest.raises(NotImplementedError):
        L.SimpleRNN(L.Input(shape=(4, 2)), 3, unroll=True)
    r = L.Reshape(L.Input(shape=(2, 3)), shape=(-1))
    assert r.oshape == (6,)
    # Q3: a bias_initializer argument selects the KERNEL initializer
    d = L.Dense(L.Input(shape=(3,)), 2, kernel_initializer=ns.init.Ones(), bias_initializer=ns.init.Zeros())
    d.initialize()
    assert float(np.asarray(d.params["b1"]).sum()) == 2.0


def test_save_load_roundtrip_host(tmp_path, ns):
    from tests import topologies
    np.random.seed(1)
    net = topologies.iris_mlp(ns)
    f = str(tmp_path / "m.pkl")
    net.save(f)
    data = pickle.load(open(f, "rb"))
    assert all(isinstance(v, np.ndarray) for d in data["params"] for v in d.values())
    before = net.get_params()
    np.random.seed(2)
    net2 = topologies.iris_mlp(ns)
    net2.load(f)
    for a, b in zip(before, net2.get_params()):
        for k in a:
            assert np.array_equal(a[k], b[k])


def test_shard_bounds():
    from dnpy_b200 import parallel
    cover = []
    for r in range(4):
        lo, hi = parallel.shard_bounds(1024, rank=r, world=4)
        cover += list(range(lo, hi))
    assert cover == list(range(1024))
    assert parallel.shard_bounds(10, rank=2, world=3) == (6, 9)


@pytest.mark.timeout(180)
def test_two_rank_gloo_gradient_exchange(tmp_path):
    """world_size-2 CPU run of the data-parallel host logic: row sharding + flat all-reduce with the
    reference's mean-gradient scaling reproduces the single-process gradient of a Dense layer."""
    script = tmp_path / "dp.py"
    script.write_text(textwrap.dedent(f"""
        import sys
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch
        from dnpy_b200 import parallel
        from dnpy_b200.runtime import config
        from oracle import dnpy_oracle as O
        rank, world = parallel.init(backend="gloo")
        rs = np.random.RandomState(0)
        x, w, b, d = rs.randn(8, 5), rs.randn(5, 3), rs.randn(1, 3), rs.randn(8, 3)
        _, gw_full, gb_full = O.dense_backward(x, w, b, d, (0.0, 0.02), (0.01, 0.0))
        lo, hi = parallel.shard_bounds(8)
        xs, ds = parallel.shard_rows(x), parallel.shard_rows(d)
        assert xs.shape[0] == 4 and (lo, hi) == (rank * 4, rank * 4 + 4)
        # what the kernels compute per rank: local sums scaled by 1/(B_local*world), reg scaled by 1/world
        inv_m = config.inv_world / xs.shape[0]
        gw = (xs.T @ ds + config.reg_scale * O.reg_backward(w, 0.0, 0.02)) * 1.0
        gw = (xs.T @ ds) * inv_m + O.reg_backward(w, 0.0, 0.02) * config.reg_scale / 8
        gb = ds.sum(0, keepdims=True) * inv_m + O.reg_backward(b, 0.01, 0.0) * config.reg_scale / 8
        flat = torch.from_numpy(np.concatenate([gw.ravel(), gb.ravel()]))
        parallel.allreduce_flat(flat)
        want = np.concatenate([gw_full.ravel(), gb_full.ravel()])
        assert np.allclose(flat.numpy(), want, rtol=1e-12, atol=1e-12), (flat.numpy() - want)
        parallel.shutdown()
        print("rank", rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, env=env, timeout=170)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


@pytest.mark.timeout(180)
def test_two_rank_gloo_exact_exchanges(tmp_path):
    """world_size-2 CPU run of the host logic that makes data parallelism EXACT: bucketed gradient all-reduce in
    arena order, BatchNorm (mean, M2) slots merged with Chan's formula, Embedding mean (SUM) / last position (MAX), and the
    global Dropout draw split into row blocks."""
    script = tmp_path / "dp2.py"
    script.write_text(textwrap.dedent(f"""
        import sys, types
        sys.path.insert(0, {ROOT!r})
        import numpy as np, torch
        from dnpy_b200 import parallel
        from dnpy_b200.runtime import config
        from dnpy_b200.layers.regularization import _global_draw
        rank, world = parallel.init(backend="gloo")
        rs = np.random.RandomState(0)
        # 1. gradient buckets: three 'layers' own arena ranges [0,40) [40,48) [48,64); min bucket 32 floats
        G = torch.full((64,), float(rank + 1))
        layers = [object(), object(), object()]
        slots = [(layers[0], 'w'), (layers[0], 'b'), (layers[1], 'w'), (layers[2], 'w')]
        arena = dict(G=G, slots=slots, offsets=[0, 32, 40, 48], total=64)
        fake = types.SimpleNamespace(_bind_arenas=lambda: arena)
        bk = parallel.GradBuckets(fake, min_bytes=128)
        bk.layer_done(layers[0]); assert bk.sent == 40          # 40 floats >= 32: sent
        bk.layer_done(layers[1]); assert bk.sent == 40          # 8 floats: merged into the next bucket
        bk.layer_done(layers[2]); assert bk.sent == 40          # 24 floats still under the minimum
        bk.finish(); assert bk.sent == 64
        assert torch.all(G == 3.0), G
        # 2. BatchNorm statistics: slots + Chan merge == statistics of the global batch
        x = rs.randn(12, 5) * 3 + 1
        xs = parallel.shard_rows(x)
        st = torch.zeros(world, 2, 5, dtype=torch.float64)
        st[rank, 0] = torch.from_numpy(xs.mean(0)); st[rank, 1] = torch.from_numpy(((xs - xs.mean(0)) ** 2).sum(0))
        parallel.gather_slots(st)
        n, mu, M2 = len(xs), st[0, 0].numpy().copy(), st[0, 1].numpy().copy()
        for r in range(1, world):
            d = st[r, 0].numpy() - mu; nt = n + len(xs)
            mu = mu + d * len(xs) / nt; M2 = M2 + st[r, 1].numpy() + d * d * n * len(xs) / nt; n = nt
        assert np.allclose(mu, x.mean(0)) and np.allclose(M2 / n, x.var(0))
        # 3. Embedding: last GLOBAL position per vocabulary row (MAX), batch mean of the delta (SUM of local parts)
        idx = rs.randint(0, 7, (4, 3)); d = rs.randn(4, 3, 2)
        lo, hi = parallel.shard_bounds(4)
        last = torch.full((7,), -1, dtype=torch.int32)
        for b in range(lo, hi):
            for l in range(3):
                last[idx[b, l]] = max(int(last[idx[b, l]]), b * 3 + l)
        mean = torch.from_numpy(d[lo:hi].sum(0) / 4.0)
        parallel.allreduce(last, "max"); parallel.allreduce(mean, "sum")
        want_last = np.full(7, -1)
        for i, v in enumerate(idx.ravel()):
            want_last[v] = i
        assert np.array_equal(last.numpy(), want_last) and np.allclose(mean.numpy(), d.mean(0))
        # 4. the global Dropout draw: this rank's row block of the single-process draw
        np.random.seed(5); full = np.random.random((6, 4))
        np.random.seed(5); mine = _global_draw(np.random.random, (3, 4))
        assert np.array_equal(mine, full[rank * 3:(rank + 1) * 3])
        assert config.route == 1                               # the literal pool routing does not shard: per-sample in effect
        parallel.shutdown()
        print("rank", rank, "ok")
    """))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29534")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                       capture_output=True, text=True, env=env, timeout=170)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


def test_evaluate_chunking_policy_and_resident_slices(ns, monkeypatch):
    """Host logic of SURVEY 8f-1/2 that needs no device: which (loss, metric) sets may be evaluated in chunks, the
    resident-slice view protocol, and that without a device the datasets are left on the host."""
    import torch
    from dnpy_b200.net import _ResidentSlice
    from tests import topologies
    np.random.seed(0)
    net = topologies.mnist_mlp(ns, h1=8, h2=4)
    net.set_mode("test")
    assert net._can_evaluate_in_chunks(1) and net._can_evaluate_in_chunks(128)
    assert not net._can_evaluate_in_chunks(net.EVAL_CHUNK)          # nothing to gain: already one forward per chunk
    monkeypatch.setenv("DNPY_B200_EVAL_LITERAL", "1")
    assert not net._can_evaluate_in_chunks(1)
    monkeypatch.delenv("DNPY_B200_EVAL_LITERAL")
    net.metrics = [[ns.metrics.RMSE()]]                             # sqrt(mean): not a mean over samples
    assert not net._can_evaluate_in_chunks(1)
    net.metrics = [[ns.metrics.CategoricalAccuracy(), ns.metrics.MAE()]]
    assert net._can_evaluate_in_chunks(1)
    net.set_mode("train")
    assert not net._can_evaluate_in_chunks(1)

    s = _ResidentSlice(torch.arange(24, dtype=torch.float32).reshape(6, 4))
    assert len(s) == 6 and s.shape == (6, 4)
    sub = s[2:5]
    assert isinstance(sub, _ResidentSlice) and sub.shape == (3, 4) and float(sub.t[0, 0]) == 8.0
    if not torch.cuda.is_available():
        x, y = [np.zeros((4, 784))], [np.zeros((4, 10))]
        rx, ry = net._make_resident(x, y)
        assert rx is x and ry is y
