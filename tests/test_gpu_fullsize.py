"""BASELINE.json configurations at their FULL sizes, checked through size-independent properties (the oracle only
finishes small cases in seconds): Parseval / linearity for the FFT, block-size invariance and sample-for-sample
agreement with the oracle on a few channels for the 8000-channel chain, decimation identities for the FIR."""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def test_cfg4_fft_4096_quarter_gigasample_parseval_and_linearity():
    """gr_fft_vcc 4096 Blackman-Harris over 61 035 vectors (250 M samples, 2 GB; the 1 G-sample config is four
    such launches): Parseval per vector, linearity, and agreement with numpy on sampled vectors."""
    import torch
    from grb200 import blocks as B
    from grb200 import firdes
    N, nvec = 4096, 61035
    w = np.asarray(firdes.window(firdes.WIN_BLACKMAN_hARRIS, N), np.float32)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.view_as_complex(torch.randn((nvec, N, 2), generator=g, device="cuda"))
    y = torch.empty_like(x)
    f = B.fft_vcc(N, True, w, False)
    f.work_device(nvec, x, y)
    torch.cuda.synchronize()
    wt = torch.from_numpy(w).cuda()
    e_in = (x.abs() ** 2 * wt ** 2).sum(dim=1, dtype=torch.float64)
    e_out = (y.abs() ** 2).sum(dim=1, dtype=torch.float64) / N
    assert float(((e_out - e_in).abs() / e_in).max()) < 1e-5          # Parseval, every vector
    for r in (0, 1, 30517, nvec - 1):                                 # sampled vectors against the float64 DFT
        ref = np.fft.fft(x[r].cpu().numpy().astype(np.complex128) * w)
        got = y[r].cpu().numpy()
        assert np.max(np.abs(got - ref)) / np.max(np.abs(ref)) < 1e-5
    # linearity: F(a x1 + x2) = a F(x1) + F(x2) on the first 4096 vectors
    n2 = 4096
    x1, x2 = x[:n2], x[n2:2 * n2]
    y12 = torch.empty_like(x1)
    f.work_device(n2, (0.5 * x1 + x2).contiguous(), y12)
    torch.cuda.synchronize()
    err = (y12 - (0.5 * y[:n2] + y[n2:2 * n2])).abs().max() / y12.abs().max()
    assert float(err) < 1e-5


def test_cfg5_full_block_chain_invariances(orc):
    """The bench.py workload at full size (12 500 rows x 8000 channels = 1e8 samples): the same block processed
    whole, in uneven pieces, and through the host-pointer entry point gives identical sync hits; three channels are
    checked sample for sample against the oracle's demod tail on the GPU's own channelizer output."""
    import torch
    import bench
    from grb200 import chain, synth_torch
    R, M = 12500, bench.M
    dev = torch.device("cuda", 0)
    cfg = bench.chain_config(R)
    a = chain.DmrChain(cfg)
    Th = a.history_rows()
    x, active = synth_torch.wideband_block(M, R, Th, 400, 77, dev)
    torch.cuda.synchronize()      # x is produced on torch's stream; the chains below run on their own streams
    a.process_device(x, R)
    hits_a, na = a.read_hits_array()
    hits_a = np.sort(hits_a.copy(), order=["channel", "bit_index"])
    res = a.fetch()
    assert na > 1000
    # (1) uneven pieces
    b = chain.DmrChain(cfg)
    got, r0 = [], 0
    for n in (3001, 517, 4096, R - 3001 - 517 - 4096):
        b.process_device(x[r0:], n)
        h, _ = b.read_hits_array()
        got.append(h.copy())
        r0 += n
    hits_b = np.sort(np.concatenate(got), order=["channel", "bit_index"])
    assert np.array_equal(hits_a["channel"], hits_b["channel"]) and np.array_equal(hits_a["bit_index"], hits_b["bit_index"])
    # (2) host-pointer entry point (pinned staging, sub-blocks)
    c = chain.DmrChain(cfg)
    host = torch.empty((Th + R, M), dtype=torch.complex64, pin_memory=True)
    host.copy_(x)
    c.process_host(host.data_ptr(), R)
    h, _ = c.read_hits_array()
    hits_c = np.sort(h.copy(), order=["channel", "bit_index"])
    assert np.array_equal(hits_a["channel"], hits_c["channel"]) and np.array_equal(hits_a["bit_index"], hits_c["bit_index"])
    # (3) oracle tail on three channels of the GPU's channelizer output (stage isolated: bit exact)
    chans = [int(active[0]), int(active[len(active) // 2]), 4321]
    for cidx in chans:
        ycol = res["channels"][:, cidx]
        d = orc.quadrature_demod_cf(cfg.quad_gain, ycol)
        f = orc.fir_fff(cfg.rrc_taps, 1, d, order=orc.ORDER_SSE)
        m, _ = orc.mm_work(orc.mm_new(cfg.omega, cfg.gain_omega, cfg.mu, cfg.gain_mu, cfg.omega_relative_limit), f,
                           order=orc.ORDER_SSE)
        k = int(res["counts"][cidx])
        assert len(m) - 8 <= k <= len(m)
        assert np.array_equal(res["soft"][:k, cidx], m[:k])
        s = orc.slicer4(m[:k], cfg.slicer_alpha)
        assert np.array_equal(res["symbols"][:k, cidx], s)
        bits = orc.unpack_k_bits_bb(2, orc.map_bb(cfg.symbol_map, s))
        cb = orc.corr_work(orc.corr_new(cfg.access_code, cfg.threshold), bits)
        want = np.nonzero(cb & 2)[0]
        mine = hits_a["bit_index"][hits_a["channel"] == cidx]
        assert np.array_equal(np.sort(mine), want)


def test_cfg1_fir_ccf_full_size_against_oracle_samples(orc):
    """fir_filter_ccf 64 taps decimate-by-4 on 10 M samples: every output of three 4096-sample windows against the
    oracle (head, middle, tail) plus the DC-gain identity on a constant stream."""
    import torch
    from grb200 import blocks as B
    from grb200 import firdes
    n, D = 10_000_000, 4
    taps = np.resize(np.asarray(firdes.low_pass(1.0, 1.0, 0.1, 0.058), np.float32), 64).astype(np.float32)
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.view_as_complex(torch.rand((n + 63, 2), generator=g, device="cuda") * 2 - 1)
    x[:63] = 0                                        # history
    y = torch.empty(n // D, dtype=torch.complex64, device="cuda")
    blk = B.fir_filter_ccf(D, taps)
    blk.work_device(n // D, x, y)
    torch.cuda.synchronize()
    xh = x.cpu().numpy()
    yh = y.cpu().numpy()
    for o0 in (0, 1_234_567, n // D - 4096):
        seg = xh[o0 * D: o0 * D + 63 + 4096 * D]      # history-prefixed window producing outputs o0 .. o0+4095
        want = orc.fir_ccf(taps, D, seg, hist_prefixed=True)[:4096]
        got = yh[o0:o0 + 4096]
        assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 1e-5
    ones = torch.ones(n // 8 + 63, dtype=torch.complex64, device="cuda")
    y1 = torch.empty(n // 32, dtype=torch.complex64, device="cuda")
    blk.work_device(n // 32, ones, y1)
    torch.cuda.synchronize()
    assert float((y1 - complex(float(taps.sum()), 0)).abs().max()) < 1e-5
