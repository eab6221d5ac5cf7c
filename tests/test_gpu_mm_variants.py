"""Every build of the clock-recovery kernel (grcuda_clock_recovery_mm_ff_set_kernel_variant: round-1 kernel, shortest
chain, fewest instructions, TMA loader, quad ring) against the oracle, BIT EXACT: soft symbols, slicer decisions,
symbol counts, final loop state -- on one call, on a stream cut into calls of awkward sizes, and on a block that starts
hundreds of rows into its buffer (a time shard's halo: the quad ring's staging window starts beyond its first lap).
Also: two FIR plans with different tiles alive at once (the kernel's dynamic shared-memory attribute is per kernel)."""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

ARGS = (12500.0 / 4800.0, 0.25 * 0.175 ** 2, 0.5, 0.175, 0.005)
VARIANTS = [-1, 0, 3, 10, 11, 13, 16, 19, 20, 21]


def make_input(rng, n, nchan):
    from grb200 import synth
    sps = ARGS[0]
    x = np.zeros((n, nchan), np.float32)
    for c in range(nchan):
        if c % 3 == 0:      # noise only: large timing errors, steps of 0 .. 5 rows
            x[:, c] = 2.0 * rng.standard_normal(n)
        else:
            sym = rng.integers(0, 4, int(n / sps) + 2) * 2 - 3
            x[:, c] = synth.shape_symbols(sym, sps, nsamples=n) + 0.1 * rng.standard_normal(n)
    x[:, 5] = 0.0            # silence: every product is +-0
    return x


def run_variant(B, torch, x, variant, order, chunks, abs0=0, skip=0):
    n, nchan = x.shape
    blk = B.clock_recovery_mm_ff(*ARGS, nchan=nchan, order=order)
    blk.set_slicer(4, 0.0)
    blk.set_kernel_variant(variant)
    d_in = torch.from_numpy(x).cuda()
    max_out = n
    outs = [[] for _ in range(nchan)]
    sls = [[] for _ in range(nchan)]
    pos = 0
    for ch in chunks:
        # rows [pos - 16, pos + ch) of the stream: 16 rows of look-back in front of every call but the first
        lo = max(0, pos - 16)
        rows = pos + ch - lo
        d_out = torch.zeros((max_out, nchan), dtype=torch.float32, device="cuda")
        d_sl = torch.zeros((max_out, nchan), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(nchan, dtype=torch.int32, device="cuda")
        blk.work_device(rows, abs0 + lo, d_in[lo:], d_out, d_sl, max_out, d_cnt)
        torch.cuda.synchronize()
        o, s, cnt = d_out.cpu().numpy(), d_sl.cpu().numpy(), d_cnt.cpu().numpy()
        for c in range(nchan):
            outs[c].append(o[:cnt[c], c].copy())
            sls[c].append(s[:cnt[c], c].copy())
        pos += ch
    st = [blk._state(c) for c in range(nchan)]
    cn = blk.counters()
    return [np.concatenate(v) for v in outs], [np.concatenate(v) for v in sls], st, cn


@pytest.mark.parametrize("order_name", ["sse", "generic"])
def test_every_kernel_variant_is_bit_exact(orc, order_name):
    import torch
    from grb200 import blocks as B
    order = B.ORDER_SSE if order_name == "sse" else B.ORDER_GENERIC
    rng = np.random.default_rng(21)
    n, nchan = 2600, 72
    x = make_input(rng, n, nchan)
    want = []
    for c in range(nchan):
        w, _ = orc.mm_work(orc.mm_new(*ARGS), x[:, c], order=order)
        want.append(w)
    variants = VARIANTS if order_name == "sse" else [-1, 0, 11, 21]
    for v in variants:
        outs, sls, st, cn = run_variant(B, torch, x, v, order, [n])
        assert cn == {"clamped": 0, "overflow": 0}, (v, cn)
        for c in range(nchan):
            assert len(outs[c]) == len(want[c]), (v, c, len(outs[c]), len(want[c]))
            assert np.array_equal(outs[c].view(np.uint32), want[c].view(np.uint32)), (v, c)
            assert np.array_equal(sls[c], orc.slicer4(want[c], 0.0)), (v, c)


def test_variants_agree_on_chunked_stream_and_deep_start():
    import torch
    from grb200 import blocks as B
    rng = np.random.default_rng(22)
    n, nchan = 3000, 64
    x = make_input(rng, n, nchan)
    chunks = [700, 129, 64, 900, 523, 684]
    ref = run_variant(B, torch, x, 0, B.ORDER_SSE, chunks, abs0=12345)
    one = run_variant(B, torch, x, 0, B.ORDER_SSE, [n], abs0=12345)
    for c in range(nchan):   # chunked == one call, up to the symbols whose 8-row look-ahead the last call did not have
        k = min(len(ref[0][c]), len(one[0][c]))
        assert k >= len(one[0][c]) - 4 and np.array_equal(ref[0][c][:k].view(np.uint32), one[0][c][:k].view(np.uint32)), c
    for v in [10, 11, 16, 21, -1]:
        got = run_variant(B, torch, x, v, B.ORDER_SSE, chunks, abs0=12345)
        for c in range(nchan):
            assert np.array_equal(got[0][c].view(np.uint32), ref[0][c].view(np.uint32)), (v, c)
            assert np.array_equal(got[1][c], ref[1][c]), (v, c)
        assert got[2] == ref[2], v


def test_block_that_starts_deep_inside_its_buffer():
    """The loop state says 'next row = abs 0' but the buffer starts 400 rows earlier (what a time shard's halo looks
    like): every kernel must skip the same rows.  Regression: the quad ring's issue window started at group 0."""
    import torch
    from grb200 import blocks as B
    rng = np.random.default_rng(23)
    n, nchan, lead = 1800, 64, 400
    x = make_input(rng, n + lead, nchan)
    res = {}
    for v in [0, 11, 21]:
        blk = B.clock_recovery_mm_ff(*ARGS, nchan=nchan)
        blk.set_slicer(4, 0.0)
        blk.set_kernel_variant(v)
        d_in = torch.from_numpy(x).cuda()
        d_out = torch.zeros((n, nchan), dtype=torch.float32, device="cuda")
        d_sl = torch.zeros((n, nchan), dtype=torch.uint8, device="cuda")
        d_cnt = torch.zeros(nchan, dtype=torch.int32, device="cuda")
        blk.work_device(n + lead, -lead, d_in, d_out, d_sl, n, d_cnt)   # row 0 of the buffer is absolute row -lead
        torch.cuda.synchronize()
        cnt = d_cnt.cpu().numpy()
        o = d_out.cpu().numpy()
        res[v] = (cnt, np.where(np.arange(n)[:, None] < cnt[None, :], o.view(np.uint32), 0))
    for v in [11, 21]:
        assert np.array_equal(res[v][0], res[0][0]) and np.array_equal(res[v][1], res[0][1]), v


def test_two_fir_plans_alive_with_different_tiles():
    """the dynamic shared-memory attribute belongs to the kernel, not to the plan."""
    import torch
    from grb200 import blocks as B
    import orc
    rng = np.random.default_rng(24)
    x = (rng.standard_normal(40000) + 1j * rng.standard_normal(40000)).astype(np.complex64)
    big = B.fir_filter_ccf(8, rng.standard_normal(200).astype(np.float32))     # large tile (decimation 8)
    small = B.fir_filter_ccf(1, rng.standard_normal(9).astype(np.float32))     # created later, small tile
    for _ in range(2):
        for blk, D, taps_n in ((big, 8, 200), (small, 1, 9)):
            nout = (len(x) - taps_n + 1) // D
            y = blk.work(nout, x)
            assert len(y) == nout
    yb = big.work((len(x) - 199) // 8, x)
    assert np.all(np.isfinite(yb.view(np.float32)))
