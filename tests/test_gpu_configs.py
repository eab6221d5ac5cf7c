"""BASELINE.json configs at their STATED sizes through the GPU blocks, against the oracle (parity gaps):

  cfg2  single-channel DMR 4FSK chain at 48 kS/s, 60 s = 2.88 M samples, ~288 k symbols: every soft symbol, dibit,
        correlator byte and sync hit bit exact; symbols/s and the mismatch count are printed
  cfg3  160-channel PFB (16 taps/branch) + batched demod over >= 10 s (125 000 rows = 20 M samples)
  cfg5  the full 8000-channel block: the oracle's demod tail on 64 channels of the GPU's own channelizer output
        (oracle in threads), all bit exact, plus the error counters of the loop at zero
"""
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def test_cfg2_single_channel_60_seconds_bit_exact(orc):
    import torch
    from grb200 import blocks as B
    from grb200 import firdes, synth
    fs, secs = 48000.0, 60.0
    sps = fs / synth.SYMBOL_RATE                      # 10 samples per symbol
    rng = np.random.default_rng(3)
    nslots = int(secs * synth.SYMBOL_RATE / 144)
    x, sym, starts = synth.dmr_channel_baseband(rng, nslots, fs, snr_db=20.0)
    n = int(secs * fs)
    x = x[:n]
    assert len(x) == n == 2_880_000
    gain = fs / (2 * np.pi * synth.DEVIATION_HZ)
    rrc = firdes.root_raised_cosine(1.0, fs, synth.SYMBOL_RATE, synth.RRC_ALPHA, 11 * 10 + 1)
    mmargs = (sps, 0.25 * 0.175 ** 2, 0.5, 0.175, 0.005)
    code = synth.access_code_string(synth.DMR_BS_DATA_SYNC_BITS)
    # ---- oracle -------------------------------------------------------------------------------------------------
    d = orc.quadrature_demod_cf(gain, x)
    f = orc.fir_fff(rrc, 1, d, order=orc.ORDER_SSE)
    m, _ = orc.mm_work(orc.mm_new(*mmargs), f, order=orc.ORDER_SSE)
    s = orc.slicer4(m, 0.0)
    bits = orc.unpack_k_bits_bb(2, orc.map_bb(synth.SLICER_TO_DIBIT_MAP, s))
    cb = orc.corr_work(orc.corr_new(code, 2), bits)
    # ---- GPU blocks through the C ABI (host pointers, scheduler-style chunks) -------------------------------------
    t0 = time.perf_counter()
    q = B.quadrature_demod_cf(gain)
    fir = B.fir_filter_fff(1, rrc)
    mm = B.clock_recovery_mm_ff(*mmargs)
    sl = B.pager_slicer_fb(0.0)
    corr = B.correlate_access_code_bb(code, 2)
    gd = B.run(q, x, chunk=400_000)
    assert np.array_equal(gd.view(np.uint32), d.view(np.uint32))
    gf = B.run(fir, gd, chunk=333_333)
    assert np.array_equal(gf.view(np.uint32), f.view(np.uint32))
    # digital_clock_recovery_mm_ff is a gr_block, not a sync block: scheduler-style calls, unconsumed items re-presented
    pos, outs = 0, []
    while pos < len(gf) - 8:
        y2, c2 = mm.general_work(30_000, gf[pos: pos + 250_000], abs_index0=pos)
        if c2 == 0 and len(y2) == 0:
            break
        outs.append(y2)
        pos += c2
    gm = np.concatenate(outs)
    k = len(gm)
    assert len(m) - 8 <= k <= len(m)
    assert np.array_equal(gm.view(np.uint32), m[:k].view(np.uint32))
    gs = B.run(sl, gm, chunk=100_000)
    gbits = orc.unpack_k_bits_bb(2, orc.map_bb(synth.SLICER_TO_DIBIT_MAP, gs))
    gc = B.run(corr, gbits, chunk=77_777)
    dt = time.perf_counter() - t0
    mism = int(np.sum(gs != s[:k])) + int(np.sum(gc != cb[:2 * k]))
    hits = int(np.sum(gc & 2))
    print("cfg2: %d samples, %d symbols, %d sync hits, %d mismatches, %.0f symbols/s through the host-pointer blocks"
          % (n, k, hits, mism, k / dt))
    assert mism == 0
    assert k > 287_000 and hits >= int(0.9 * len(starts))
    assert mm.counters() == {"clamped": 0, "overflow": 0}


def test_cfg3_ten_seconds_160_channels(orc):
    import torch
    from test_gpu_chain import make_cfg, oracle_tail
    from grb200 import chain, synth_torch
    M, T, rows = 160, 16, 125_000                      # 10 s at 12.5 kS/s per channel = 20 M wideband samples
    dev = torch.device("cuda", 0)
    cfg = make_cfg(M, T, max_rows=rows, keep_bytes=False)
    ch = chain.DmrChain(cfg)
    Th = ch.history_rows()
    x, active = synth_torch.wideband_block(M, rows, Th, 16, 4, dev, noise_sigma=2e-3)
    x[:Th] = 0
    torch.cuda.synchronize()
    s0 = torch.cuda.current_stream().cuda_stream
    ch.process_front_device(x, rows, s0)
    ch.process_tail_device(s0)
    res = ch.fetch()
    hits, nh = ch.read_hits_array()
    assert ch.counters() == {"clamped": 0, "overflow": 0, "hits_dropped": 0}
    # channelizer against the oracle on the first 4096 rows (the oracle's float64 DFT is slow), then the tail of 12
    # channels over the whole 10 s on the GPU's own channelizer output
    xs = x[Th:Th + 4096].cpu().numpy().reshape(-1)
    want, _ = orc.pfb_channelizer_ccf(M, cfg.pfb_taps, xs)
    got = res["channels"][:want.shape[0]]
    assert float(np.max(np.abs(got - want)) / np.max(np.abs(want))) < 1e-5
    chans = [int(c) for c in active[:8]] + [7, 77, 133, 159]

    def one(c):
        m, s, cb = oracle_tail(orc, cfg, res["channels"][:, c], orc.ORDER_SSE)
        k = int(res["counts"][c])
        assert len(m) - 8 <= k <= len(m), (c, k, len(m))
        assert np.array_equal(res["soft"][:k, c].view(np.uint32), m[:k].view(np.uint32)), c
        assert np.array_equal(res["symbols"][:k, c], s[:k]), c
        want_hits = np.nonzero(cb[:2 * k] & 2)[0]
        mine = np.sort(hits["bit_index"][hits["channel"] == c])
        assert np.array_equal(mine, want_hits), c
        return len(want_hits)
    with ThreadPoolExecutor(8) as ex:
        nhits = list(ex.map(one, chans))
    assert sum(nhits[:8]) >= 8 * 300                  # ~33 bursts per second per active channel


def test_cfg5_full_block_oracle_tail_on_64_channels(orc):
    import torch
    import bench
    from grb200 import chain, synth_torch
    R, M = 12500, bench.M
    dev = torch.device("cuda", 0)
    cfg = bench.chain_config(R)
    a = chain.DmrChain(cfg)
    Th = a.history_rows()
    x, active = synth_torch.wideband_block(M, R, Th, 800, 78, dev)
    torch.cuda.synchronize()
    for split in (True, False):                       # the two-kernel tail and the fused one
        ch = chain.DmrChain(cfg)
        ch.set_split_correlator(split)
        s0 = torch.cuda.current_stream().cuda_stream
        ch.process_front_device(x, R, s0)
        ch.process_tail_device(s0)
        res = ch.fetch()
        hits, nh = ch.read_hits_array()
        assert ch.counters() == {"clamped": 0, "overflow": 0, "hits_dropped": 0}
        if split:
            first = (res["counts"].copy(), res["soft"].copy(), np.sort(hits.copy(), order=["channel", "bit_index"]))
        else:
            h2 = np.sort(hits.copy(), order=["channel", "bit_index"])
            assert np.array_equal(first[0], res["counts"])
            assert np.array_equal(first[2]["channel"], h2["channel"]) and np.array_equal(first[2]["bit_index"], h2["bit_index"])
        del ch
    rng = np.random.default_rng(0)
    chans = sorted(set([int(c) for c in active[::25]] + [int(c) for c in rng.choice(M, 32, replace=False)]))[:64]
    assert len(chans) >= 60
    ycols = {c: np.ascontiguousarray(res["channels"][:, c]) for c in chans}

    def one(c):
        d = orc.quadrature_demod_cf(cfg.quad_gain, ycols[c])
        f = orc.fir_fff(cfg.rrc_taps, 1, d, order=orc.ORDER_SSE)
        m, _ = orc.mm_work(orc.mm_new(cfg.omega, cfg.gain_omega, cfg.mu, cfg.gain_mu, cfg.omega_relative_limit), f, order=orc.ORDER_SSE)
        k = int(res["counts"][c])
        assert len(m) - 8 <= k <= len(m), (c, k, len(m))
        assert np.array_equal(res["soft"][:k, c].view(np.uint32), m[:k].view(np.uint32)), c
        s = orc.slicer4(m[:k], cfg.slicer_alpha)
        assert np.array_equal(res["symbols"][:k, c], s), c
        bits = orc.unpack_k_bits_bb(2, orc.map_bb(cfg.symbol_map, s))
        cb = orc.corr_work(orc.corr_new(cfg.access_code, cfg.threshold), bits)
        want = np.nonzero(cb & 2)[0]
        mine = np.sort(hits["bit_index"][hits["channel"] == c])
        assert np.array_equal(mine, want), c
        return len(want)
    with ThreadPoolExecutor(8) as ex:
        tot = sum(ex.map(one, chans))
    assert tot > 200
