"""GPU parity of the blocks either side of the hot path (SURVEY 8f rank 4 + the stand-alone forms of a12 / a14):
gr_framer_sink_1, digital_clock_recovery_mm_cc, gr_map_bb, gr_unpack_k_bits_bb, gr_stream_to_streams,
gr_vector_to_streams -- through the C ABI (grb200.blocks is a ctypes mirror) against the oracle and the committed
fixtures of the compiled reference.  Byte / index work and the feedback loop: bit exact."""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


@pytest.fixture(scope="module")
def B():
    from grb200 import blocks
    return blocks


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


# ---- gr_map_bb / gr_unpack_k_bits_bb ---------------------------------------------------------------------------------
@pytest.mark.parametrize("nmap,n", [(4, 1), (4, 15), (4, 4099), (0, 1000), (256, 70001), (300, 5000)])
def test_map_bb(B, orc, nmap, n):
    rng = np.random.default_rng(nmap + n)
    m = rng.integers(0, 256, nmap)
    x = rng.integers(0, 256, n).astype(np.uint8)
    want = orc.map_bb(m, x)
    assert np.array_equal(B.map_bb(m).work(n, x), want)
    tab = np.arange(256)
    tab[:min(nmap, 256)] = m[:256]
    assert np.array_equal(want, tab[x].astype(np.uint8))          # gr_map_bb.cc:40-45, 58-59


def test_map_bb_device_unaligned(B):
    import torch
    m = [3, 2, 1, 0]
    x = torch.randint(0, 4, (100003,), dtype=torch.uint8, device="cuda")
    y = torch.zeros_like(x)
    blk = B.map_bb(m)
    blk.work_device(100000, x[3:], y[3:])                          # 16-byte path off: byte loop
    blk.work_device(2, x, y)
    torch.cuda.synchronize()
    assert torch.equal(y[3:100003], 3 - x[3:100003]) and torch.equal(y[:2], 3 - x[:2]) and int(y[2]) == 0


@pytest.mark.parametrize("k,n", [(1, 100), (2, 4801), (3, 1000), (8, 4096), (12, 77), (32, 64), (40, 10)])
def test_unpack_k_bits(B, orc, k, n):
    rng = np.random.default_rng(k)
    x = rng.integers(0, 256, n).astype(np.uint8)
    blk = B.unpack_k_bits_bb(k)
    assert blk.interpolation() == k
    y = blk.work(n * k, x)
    if k <= 32:
        assert np.array_equal(y, orc.unpack_k_bits_bb(k, x))
    want = np.array([(int(t) >> j) & 1 if j < 32 else 0 for t in x for j in range(k - 1, -1, -1)], np.uint8)
    assert np.array_equal(y, want)
    assert len(blk.work(n * k + k - 1, x)) == n * k                # noutput_items / k input bytes are read (:60)


def test_unpack_k_bits_errors(B):
    with pytest.raises(IndexError):
        B.unpack_k_bits_bb(0)                                      # std::out_of_range (:44-45)


# ---- gr_stream_to_streams / gr_vector_to_streams ---------------------------------------------------------------------
@pytest.mark.parametrize("dtype,ns,n", [(np.uint8, 3, 1000), (np.int16, 7, 513), (np.float32, 8, 4097), (np.complex64, 160, 300),
                                        (np.complex64, 8000, 70), (np.complex128, 5, 100), (np.complex64, 1, 50)])
def test_stream_to_streams(B, dtype, ns, n):
    rng = np.random.default_rng(ns)
    x = rng.integers(0, 255, n * ns * np.dtype(dtype).itemsize).astype(np.uint8).view(dtype)
    for cls in (B.stream_to_streams, B.vector_to_streams):
        outs = cls(np.dtype(dtype).itemsize, ns).work(n, x)
        assert len(outs) == ns
        want = x.reshape(n, ns)
        for j in range(ns):
            assert np.array_equal(outs[j].view(np.uint8), np.ascontiguousarray(want[:, j]).view(np.uint8))


def test_stream_to_streams_odd_item_size(B):
    x = np.arange(12 * 5 * 9, dtype=np.uint8)                      # 12-byte items, 5 streams
    outs = B.stream_to_streams(12, 5).work(9, x)
    want = x.reshape(9, 5, 12)
    for j in range(5):
        assert np.array_equal(outs[j], want[:, j].reshape(-1))
    with pytest.raises(ValueError):
        B.stream_to_streams(0, 4)


def test_streams_feed_the_channelizer_entry_points(B, orc):
    """blks2.pfb_channelizer_ccf (blks2impl/pfb_channelizer.py:61-75): stream_to_streams in front of the channelizer's
    M inputs == the interleaved entry point."""
    M, rows = 20, 200
    rng = np.random.default_rng(3)
    taps = rng.standard_normal(M * 6).astype(np.float32)
    x = (rng.standard_normal(M * rows) + 1j * rng.standard_normal(M * rows)).astype(np.complex64)
    blk = B.pfb_channelizer_ccf(M, taps)
    T = blk.taps_per_filter()
    full = np.concatenate([np.zeros(T * M, np.complex64), x])
    streams = B.stream_to_streams(8, M).work(T + rows, full)
    blk.general_work(rows, streams)
    y, c = blk.general_work(rows, streams)
    want, wc = orc.pfb_channelizer_ccf(M, taps, x)
    assert c == wc and np.max(np.abs(y - want)) / np.max(np.abs(want)) < 5e-6


# ---- gr_framer_sink_1 ------------------------------------------------------------------------------------------------
def split_packets(fx):
    out, pos = [], 0
    for off, n in zip(fx["framer_offsets"], fx["framer_lengths"]):
        out.append((int(off), bytes(fx["framer_payloads"][pos:pos + n])))
        pos += n
    return out


@pytest.mark.parametrize("nchunks", [1, 23, 997])
def test_framer_fixture_single_stream(B, golden_next, nchunks):
    want = split_packets(golden_next)
    f, got = B.framer_sink_1(), []
    for chunk in np.array_split(golden_next["framer_stream"], nchunks):
        assert f.work(chunk) == len(chunk)
        got += [(o, p) for _, o, p in f.messages()]
    assert got == want and f.dropped == 0


def make_streams(orc, rng, nchan, corrupt=3, maxlen=300):
    streams, want = [], []
    for c in range(nchan):
        pk = [(int(rng.integers(0, 16)), bytes(rng.integers(0, 256, int(rng.integers(0, maxlen))).astype(np.uint8)))
              for _ in range(int(rng.integers(0, 7)))]
        s = orc.framer_make_stream(rng, pk, gap=(0, 90), corrupt_header_every=corrupt)
        s[rng.integers(0, len(s), 4)] |= 2                        # stray flags: ignored inside a packet, honoured outside
        streams.append(s)
    return streams


@pytest.mark.parametrize("layout", ["rows", "stream"])
def test_framer_batched_vs_oracle(B, orc, layout):
    """70 channels, ragged lengths, fed in 5 pieces: every channel's messages equal the oracle's on its stream, in both
    data layouts (thread per channel on [item][channel] rows, warp per channel on stream-major data)."""
    import torch
    rng = np.random.default_rng(11)
    nchan = 70
    streams = make_streams(orc, rng, nchan)
    L = max(len(s) for s in streams)
    want = {}
    for c, s in enumerate(streams):
        o = orc.Framer()
        want[c] = o.work(s)
    f = B.framer_sink_1(nchan)
    got = {c: [] for c in range(nchan)}
    cuts = sorted({0, 37, 900, 901, L // 2, L})
    for a, b in zip(cuts[:-1], cuts[1:]):
        cnt = np.array([max(0, min(len(s), b) - a) for s in streams], np.int32)
        n = int(cnt.max())
        buf = np.zeros((nchan, n), np.uint8)
        for c, s in enumerate(streams):
            buf[c, :cnt[c]] = s[a:a + cnt[c]]
        d_cnt = torch.from_numpy(cnt).cuda()
        if layout == "rows":
            d = torch.from_numpy(np.ascontiguousarray(buf.T)).cuda()         # [item][channel]
            f.work_device(n, d, nchan, 1, d_cnt, 1)
        else:
            d = torch.from_numpy(buf).cuda()                                 # [channel][item]
            f.work_device(n, d, 1, n, d_cnt, 1)
        last_end = -1
        for c, off, p, end, seq in f.messages(with_meta=True):
            assert end >= last_end                                            # arrival order of one shared queue
            last_end = end
            assert seq == len(got[c])
            got[c].append((off, p))
        assert f.dropped == 0
    assert got == want
    assert sum(len(v) for v in want.values()) > 50


def test_framer_long_payload_and_queue_limits(B, orc):
    rng = np.random.default_rng(2)
    pk = [(5, bytes(rng.integers(0, 256, 4095).astype(np.uint8))), (1, b""), (2, b"x")]
    s = orc.framer_make_stream(rng, pk)
    f = B.framer_sink_1()
    f.work(s)
    assert [(o, p) for _, o, p in f.messages()] == pk == orc.Framer().work(s)
    small = B.framer_sink_1(1, max_msgs=2, payload_capacity=64)
    small.work(s)
    m = small.messages()
    assert len(m) == 2 and m[0][2] is None and m[1] == (0, 1, b"") and small.dropped == 2   # payload lost, third message lost
    with pytest.raises(ValueError):
        B.framer_sink_1(0)


def test_framer_behind_the_chain_bytes(B, orc):
    """The correlator byte stream of the chain ([bit][channel], 2 bits per symbol) straight into the batched framer:
    per channel identical to the oracle framer on the same bytes."""
    import torch
    from grb200 import chain, firdes, synth
    M, T, rows = 160, 16, 2400
    rng = np.random.default_rng(0)
    active = [3, 42, 159]
    x, _ = synth.wideband_compose(rng, M, rows, active, noise_sigma=2e-3)
    taps = firdes.low_pass_2(1.0, M * 12500.0, 5500.0, 1500.0, 60.0, firdes.WIN_BLACKMAN_hARRIS)
    c0 = len(taps) // 2
    taps = (taps[c0 - M * T // 2: c0 - M * T // 2 + M * T] * M).astype(np.float32)
    ch = chain.DmrChain(chain.DmrChainConfig(M, taps, max_rows_per_block=rows, keep_bytes=True))
    buf = torch.from_numpy(np.concatenate([np.zeros((ch.history_rows(), M), np.complex64), x.reshape(rows, M)])).cuda()
    ch.process_device(buf, rows)
    torch.cuda.synchronize()
    res = ch.fetch()
    byts, counts = res["bytes"], res["counts"]
    # make packets likely: set the flag where the data happens to spell a valid header (it never does by chance), i.e.
    # plant headers into the byte stream of a few channels before framing
    byts = byts.copy()
    for c in (3, 42, 100):
        n = 2 * int(counts[c])
        s = orc.framer_make_stream(rng, [(7, b"DMR"), (1, bytes(range(40)))], gap=(3, 30))
        byts[100:100 + len(s), c] = s[: max(0, n - 100)][: len(s)]
    f = B.framer_sink_1(M)
    d = torch.from_numpy(byts).cuda()
    f.work_device(byts.shape[0], d, M, 1, torch.from_numpy(counts.astype(np.int32)).cuda(), 2)
    got = {}
    for c, off, p in f.messages():
        got.setdefault(c, []).append((off, p))
    for c in range(M):
        want = orc.Framer().work(byts[: 2 * int(counts[c]), c])
        assert got.get(c, []) == want, c
    assert len(got[3]) == 2 and got[42][0] == (7, b"DMR")


# ---- digital_clock_recovery_mm_cc ------------------------------------------------------------------------------------
def test_mm_cc_fixture(B, golden_next):
    fx = golden_next
    args = [float(v) for v in fx["mmcc_args"]]
    y, e, c = B.clock_recovery_mm_cc(*args).general_work(len(fx["mmcc_x"]), fx["mmcc_x"])
    assert c == int(fx["mmcc_consumed_plain"]) and np.array_equal(bits(y), bits(fx["mmcc_y_plain"])) and e is None
    y, e, c = B.clock_recovery_mm_cc(*args).general_work(len(fx["mmcc_x"]), fx["mmcc_x"], with_error=True)
    assert c == int(fx["mmcc_consumed_err"]) and np.array_equal(bits(y), bits(fx["mmcc_y_err"]))
    assert np.array_equal(bits(e), bits(fx["mmcc_err"]))


def qpsk(rng, omega, n, noise=0.1):
    sps = int(np.ceil(omega))
    sym = (rng.integers(0, 2, n) * 2 - 1) + 1j * (rng.integers(0, 2, n) * 2 - 1)
    return (np.repeat(sym, sps)[:n] + noise * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)


@pytest.mark.parametrize("omega,gm,n", [(2.0, 0.05, 4000), (4.0, 0.1, 6000), (8.0, 0.175, 8000), (1.0, 0.01, 2000), (2.6041667, 0.175, 5000)])
def test_mm_cc_scheduler_style_vs_oracle(B, orc, omega, gm, n):
    rng = np.random.default_rng(int(omega * 10))
    x = qpsk(rng, omega, n)
    for we in (False, True):
        blk = B.clock_recovery_mm_cc(omega, 0.25 * gm * gm, 0.5, gm, 0.005)
        st = orc.mmcc_new(omega, 0.25 * gm * gm, 0.5, gm, 0.005)
        assert blk.history() == 3
        assert blk.forecast(100) == orc.lib().orc_mmcc_forecast(__import__("ctypes").byref(st), 100)
        pos = 0
        while pos < n - 64:
            yo, eo, co = orc.mmcc_work(st, x[pos:], noutput=257, with_error=we)
            yg, eg, cg = blk.general_work(257, x[pos:], with_error=we)
            assert cg == co and np.array_equal(bits(yg), bits(yo))
            if we:
                assert np.array_equal(bits(eg), bits(eo))
            if co == 0:
                break
            pos += co
        mu, om = blk.state()
        assert bits(np.float32(mu)) == bits(np.float32(st.mu)) and bits(np.float32(om)) == bits(np.float32(st.omega))
        assert blk.counters()[0] == 0      # (calls that end at noutput_items are the scheduler's normal case here)


def test_mm_cc_batched_device_vs_oracle(B, orc):
    """70 channels [time][channel], three blocks with carried positions: every channel equals the oracle on its column."""
    import torch
    rng = np.random.default_rng(8)
    nchan, n, omega, gm = 70, 6000, 2.6041667, 0.175
    X = np.stack([qpsk(rng, omega, n, noise=0.05 + 0.01 * c) * np.float32(1 + 0.02 * c) for c in range(nchan)], axis=1)   # [time][channel]
    for we in (False, True):
        blk = B.clock_recovery_mm_cc(omega, 0.25 * gm * gm, 0.5, gm, 0.005, nchan=nchan)
        d = torch.from_numpy(np.ascontiguousarray(X)).cuda()
        max_out = 1200
        got = [[] for _ in range(nchan)]
        gerr = [[] for _ in range(nchan)]
        out = torch.zeros((max_out, nchan), dtype=torch.complex64, device="cuda")
        err = torch.zeros((max_out, nchan), dtype=torch.float32, device="cuda") if we else None
        cnt = torch.zeros(nchan, dtype=torch.int32, device="cuda")
        for a, b in ((0, 2000), (1900, 4100), (4000, 6000)):       # overlapping windows: each channel resumes where it stopped
            blk.work_device(b - a, a, d[a:b], out, err, max_out, cnt)
            torch.cuda.synchronize()
            o, k = out.cpu().numpy(), cnt.cpu().numpy()
            for c in range(nchan):
                got[c].append(o[:k[c], c])
                if we:
                    gerr[c].append(err.cpu().numpy()[:k[c], c])
        assert blk.counters() == (0, 0)
        for c in range(nchan):
            st = orc.mmcc_new(omega, 0.25 * gm * gm, 0.5, gm, 0.005)
            yo, eo, _ = orc.mmcc_work(st, np.ascontiguousarray(X[:, c]), with_error=we)
            y = np.concatenate(got[c])
            assert len(y) >= len(yo) - 12 and np.array_equal(bits(y[:len(yo)]), bits(yo[:len(y)])), c
            if we:
                e = np.concatenate(gerr[c])
                assert np.array_equal(bits(e[:len(eo)]), bits(eo[:len(e)]))


def test_mm_cc_constructor_errors_and_setters(B, orc):
    with pytest.raises(IndexError):
        B.clock_recovery_mm_cc(0.0, 0.1, 0.5, 0.1)
    with pytest.raises(IndexError):
        B.clock_recovery_mm_cc(2.0, -0.1, 0.5, 0.1)
    with pytest.raises(IndexError):
        B.clock_recovery_mm_cc(2.0, 0.1, 0.5, -0.1)
    rng = np.random.default_rng(5)
    x = qpsk(rng, 4.0, 3000)
    blk = B.clock_recovery_mm_cc(4.0, 0.001, 0.5, 0.05, 0.01)
    blk.set_mu(0.25)
    blk.set_omega(4.02)
    blk.set_gain_mu(0.07)
    blk.set_gain_omega(0.002)
    st = orc.mmcc_new(4.0, 0.002, 0.25, 0.07, 0.01)
    om, lim = float(np.float32(4.02)), float(np.float32(0.01))
    st.omega = om                                                  # set_omega (.h:75-80) also moves omega_mid
    mn, mx = np.float32(om * (1.0 - lim)), np.float32(om * (1.0 + lim))
    st.omega_mid = float(np.float32(0.5 * float(np.float32(mn + mx))))
    yo, _, co = orc.mmcc_work(st, x)
    yg, _, cg = blk.general_work(len(x), x)
    assert cg == co and np.array_equal(bits(yg), bits(yo))
