"""CPU checks of the ALGORITHMIC restructurings the CUDA kernels rely on (no GPU, no CUDA code run here): each test restates
in NumPy what a kernel computes differently from the reference's loop order and compares it with the float64 oracle, so a
parity failure on the GPU can be told apart from a flaw in the scheme itself.

* csrc/rnn_tc.cu — truncated BPTT as (a) a serial chain pre(t) and (b) lag recurrences over flat (sample, step) rows in
  overlapping tiles with the lag sum W(u) = sum_k S_k(u + k) and one reduction per tile (recurrent.py:190-266);
* csrc/conv.cu — fp32-grade products from three TF32 products of hi / lo split operands;
* csrc/tc_conv_c1p.cu — the bias carried as three BF16 pieces, the block-diagonal image-group GEMM over im2col'ed chunks.
"""
import numpy as np
import pytest

from oracle import dnpy_oracle as O


def _tf32_trunc(a):
    """value kept by a TF32 operand: the low 13 mantissa bits dropped"""
    return (np.asarray(a, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)


@pytest.mark.parametrize("B,T,H,D,trunc,tile", [(3, 11, 5, 3, 2, 8), (2, 40, 6, 2, 4, 16), (5, 7, 4, 3, 7, 16), (1, 3, 3, 1, 1, 4)])
def test_rnn_lag_sum_over_overlapping_row_tiles_equals_literal_bptt(B, T, H, D, trunc, tile):
    rs = np.random.RandomState(B * 100 + T)
    x = rs.randn(B, T, D)
    waa, wax, ba = rs.randn(H, H) * 0.4, rs.randn(H, D) * 0.4, rs.randn(1, H) * 0.1
    states = O.rnn_forward(x, waa, wax, ba)
    states = states[0] if isinstance(states, tuple) else states
    delta = rs.randn(B, H)
    want_dx, want_gwaa, want_gwax, want_gba = O.rnn_backward(x, states, delta, waa, wax, bptt_truncate=trunc)

    # (a) the serial chain (rnn_chain_kernel): pre(t) = (dy[t == T-1] + pre(t+1) . Waa) * (1 - h_t^2), dx_t = pre(t) . Wax
    pre = np.zeros((B, T, H))
    dh = delta.copy()
    for t in reversed(range(T)):
        pre[:, t] = dh * (1 - states[:, t] ** 2)
        dh = pre[:, t] @ waa
    assert np.allclose(pre @ wax, want_dx, rtol=1e-12, atol=1e-12)

    # (b) rnn_lags_tc_kernel: flat rows r = b*T + t, tiles of `tile` rows advancing by tile - lags (a tile owns its first
    # tile - lags rows), lag k+1 of row r adds into W(r - k - 1), one reduction per tile against h(u-1) | x(u) | 1
    lags = min(T - 1, trunc)
    R = B * T
    h_flat, x_flat, pre_flat = states.reshape(R, H), x.reshape(R, D), pre.reshape(R, H)
    step = tile - lags
    gwaa, gwax, gba = np.zeros((H, H)), np.zeros((H, D)), np.zeros((1, H))
    for r0 in range(0, R, step):
        rows = np.arange(r0, r0 + tile)
        valid = rows < R
        t_of = np.where(valid, rows % T, 0)
        S = np.where(valid[:, None], pre_flat[np.minimum(rows, R - 1)], 0.0)
        W = S.copy()                                          # lag 0
        for k in range(lags):
            live = valid & (t_of - k - 1 >= 0)
            hprev = h_flat[np.clip(rows - k - 1, 0, R - 1)]
            A = np.where(live[:, None], S * (1 - hprev ** 2), 0.0)
            S = A @ waa                                       # S_{k+1}(row)
            for i in np.nonzero(live)[0]:
                u = i - k - 1
                if u >= 0:                                    # rows before the tile belong to the previous tile
                    W[u] += S[i]
        owned = valid & (np.arange(tile) < step)
        hm1 = np.where((owned & (t_of >= 1))[:, None], h_flat[np.clip(rows - 1, 0, R - 1)], 0.0)
        xm = np.where(owned[:, None], x_flat[np.minimum(rows, R - 1)], 0.0)
        gwaa += W.T @ hm1
        gwax += W.T @ xm
        gba += (W * owned[:, None]).sum(axis=0, keepdims=True)
    assert np.allclose(gwaa, want_gwaa, rtol=1e-10, atol=1e-12)
    assert np.allclose(gwax, want_gwax, rtol=1e-10, atol=1e-12)
    assert np.allclose(gba, want_gba, rtol=1e-10, atol=1e-12)


def test_three_tf32_products_of_split_operands_are_fp32_grade():
    rs = np.random.RandomState(0)
    a = (rs.randn(64, 288) * rs.choice([1e-3, 1.0, 30.0], (64, 288))).astype(np.float32)
    b = rs.randn(288, 32).astype(np.float32)
    a_hi, b_hi = _tf32_trunc(a), _tf32_trunc(b)
    a_lo, b_lo = a - a_hi, b - b_hi                          # exact in fp32
    assert np.array_equal(a_hi.astype(np.float64) + a_lo.astype(np.float64), a.astype(np.float64))
    f = lambda m: _tf32_trunc(m).astype(np.float64)          # what the tensor core reads of an operand
    got = f(a_hi) @ f(b_hi) + f(a_lo) @ f(b_hi) + f(a_hi) @ f(b_lo)
    want = a.astype(np.float64) @ b.astype(np.float64)
    one = f(a) @ f(b)                                        # plain TF32
    err3 = np.abs(got - want).max() / np.abs(want).max()
    err1 = np.abs(one - want).max() / np.abs(want).max()
    assert err3 <= 2.0 ** -20 and err1 > 50 * err3, (err3, err1)


def test_bias_as_three_bf16_pieces_is_exact_to_fp32():
    import torch
    v = torch.from_numpy((np.random.RandomState(1).randn(4096) * 9).astype(np.float32))
    hi = v.bfloat16().float()
    r1 = v - hi
    mid = r1.bfloat16().float()
    lo = (r1 - mid).bfloat16().float()
    back = (hi.double() + mid.double() + lo.double()).numpy()
    assert np.abs(back - v.double().numpy()).max() <= np.abs(v.numpy()).max() * 2.0 ** -24


def test_block_diagonal_image_group_gemm_equals_the_convolution():
    """tc_c1p_kernel: A row (p, f) holds filter f's taps in K columns [16p, 16p + 9) and the bias in columns 16p + 9..11, B row n
    is conv pixel n = y * NC + x of a chunk with image p's 3x3 patch in the same K columns and ones behind it."""
    rs = np.random.RandomState(2)
    F, G, W_, NC, rows = 32, 4, 28, 25, 5
    x = rs.rand(G, W_, W_)
    w = rs.randn(F, 9)
    bias9 = rs.randn(F)                                       # already K * bias
    for chunk in (0, 3, 5):
        A = np.zeros((G * F, 16 * G))
        for p in range(G):
            A[p * F:(p + 1) * F, 16 * p:16 * p + 9] = w
            A[p * F:(p + 1) * F, 16 * p + 9] = bias9          # (hi + mid + lo: one column here)
        Bm = np.zeros((rows * NC, 16 * G))
        for n in range(rows * NC):
            y, xx = divmod(n, NC)
            for p in range(G):
                Bm[n, 16 * p:16 * p + 9] = x[p, 4 * chunk + y:4 * chunk + y + 3, xx:xx + 3].reshape(9)
                Bm[n, 16 * p + 9] = 1.0
        Dm = A @ Bm.T
        conv = O.conv2d_forward(x[:, None], w.reshape(F, 1, 3, 3), bias9 / 9.0, (1, 1), (0, 0, 0, 0), (F, 26, 26))
        for p in range(G):
            want = conv[p, :, 4 * chunk:4 * chunk + rows, :NC].reshape(F, rows * NC)
            assert np.allclose(Dm[p * F:(p + 1) * F], want, rtol=1e-12, atol=1e-12)


def _pool_window_bookkeeping(win, nansafe):
    """The fused kernels' 3x3 window walk (fused_cpa.cu cpa_walk, tc_conv_c1p.cu c1p_pool_chunk): per conv row the 3-max and the
    FIRST column equal to it, rows combined with a strict > (an earlier row keeps a tie).  Returns (value, offset, nan_seen)."""
    best, idx, nan_seen = None, 0, False
    for i in range(3):
        c0, c1, c2 = win[i]
        if nansafe:                                            # np.argmax's sequential comparison: first NaN wins and sticks
            hb, hj = c0, 0
            if c1 > hb or (c1 != c1 and hb == hb): hb, hj = c1, 1
            if c2 > hb or (c2 != c2 and hb == hb): hb, hj = c2, 2
        else:                                                  # max.NaN + first-equal
            hb = np.nan if (c0 != c0 or c1 != c1 or c2 != c2) else max(c0, c1, c2)
            hj = 0 if c0 == hb else (1 if c1 == hb else 2)
        if i == 0:
            best, idx = hb, hj
        else:
            take = (hb > best or (hb != hb and best == best)) if nansafe else (hb > best)
            if not nansafe:
                nan_seen = nan_seen or best != best or hb != hb
                best = np.nan if (best != best or hb != hb) else max(best, hb)
            elif take:
                best = hb
            if take:
                idx = 3 * i + hj
    return best, idx, nan_seen or (best != best)


def test_row_wise_max_with_strict_row_combination_is_numpy_argmax():
    rs = np.random.RandomState(3)
    for trial in range(4000):
        win = rs.randint(-2, 3, (3, 3)).astype(np.float64)     # many ties
        if trial % 5 == 0:
            win[rs.randint(3), rs.randint(3)] = np.nan
        if trial % 11 == 0:
            win[rs.randint(3), rs.randint(3)] = np.nan
        want = int(np.argmax(win.reshape(9)))
        val, idx, nan_seen = _pool_window_bookkeeping(win, nansafe=False)
        if np.isnan(win).any():
            assert nan_seen                                    # the fast path only DETECTS it; the kernel then repeats NANSAFE
            val, idx, _ = _pool_window_bookkeeping(win, nansafe=True)
            assert np.isnan(val)
        else:
            assert not nan_seen and val == win.reshape(9)[want]
        assert idx == want, (win, idx, want)
