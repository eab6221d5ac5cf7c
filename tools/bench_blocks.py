#!/usr/bin/env python3
"""Device-resident throughput of the single blocks at the BASELINE.json configs that are not the bench.py workload
(configs[0] fir_filter_ccf 64 taps decimate-by-4 on 10 M samples, configs[3] fft_vcc 4096 Blackman-Harris), plus
freq_xlating_fir_filter_ccf, the stand-alone discriminator and the raw pinned H2D/D2H copy rate of the box.
Prints one JSON object; algorithmic bytes per SURVEY.md section 8d; peak = MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))


def timeit(fn, reps, flush):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def main():
    import torch
    import bench
    from grb200 import blocks as B
    from grb200 import firdes, lib
    lib.load()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    peak, kind = bench.peaks()
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 256 MB > L2: written between timed launches

    def flush():
        junk.fill_(1)

    out = {"hbm_peak_GBps": peak, "peak_kind": kind, "l2": "256 MB buffer rewritten between timed launches"}
    g = torch.Generator(device=dev).manual_seed(1)

    # cfg1: fir_filter_ccf, 64 taps, decimate by 4, 10 M complex samples (scaled x8 so that one launch is ~100 us)
    taps = np.asarray(firdes.low_pass(1.0, 1.0, 0.1, 0.058)[:64], np.float32)
    taps = np.resize(taps, 64).astype(np.float32)
    for n_in, tag in ((10_000_000, "cfg1_fir_ccf_10M"), (80_000_000, "cfg1_fir_ccf_80M")):
        x = torch.view_as_complex(torch.rand((n_in + 63, 2), generator=g, device=dev) * 2 - 1)
        nout = n_in // 4
        y = torch.empty(nout, dtype=torch.complex64, device=dev)
        blk = B.fir_filter_ccf(4, taps)
        ms = timeit(lambda: blk.work_device(nout, x, y), 10, flush)
        byt = 8.0 * n_in + 8.0 * nout
        out[tag] = {"ms": ms, "MSps_in": n_in / ms / 1e3, "alg_GBps": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / peak}
        del x, y
    # freq_xlating (complex taps + rotator), same shape
    n_in = 80_000_000
    x = torch.view_as_complex(torch.rand((n_in + 63, 2), generator=g, device=dev) * 2 - 1)
    y = torch.empty(n_in // 4, dtype=torch.complex64, device=dev)
    fx = B.freq_xlating_fir_filter_ccf(4, taps, 12500.0, 100000.0)
    ms = timeit(lambda: fx.work_device(n_in // 4, x, y), 10, flush)
    byt = 8.0 * n_in + 8.0 * (n_in // 4)
    out["freq_xlating_fir_ccf_80M"] = {"ms": ms, "MSps_in": n_in / ms / 1e3, "alg_GBps": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / peak}
    del x, y
    # cfg4: fft_vcc 4096, Blackman-Harris, forward; 61 035 vectors = 250 M samples per launch (1 G samples = 4 launches)
    N, nvec = 4096, 61035
    w = np.asarray(firdes.window(firdes.WIN_BLACKMAN_hARRIS, N), np.float32)
    x = torch.view_as_complex(torch.randn((nvec * N, 2), generator=g, device=dev))
    y = torch.empty_like(x)
    for shift in (False, True):
        f = B.fft_vcc(N, True, w, shift)
        ms = timeit(lambda: f.work_device(nvec, x, y), 10, flush)
        byt = 16.0 * nvec * N
        out["cfg4_fft_vcc_4096_bh%s" % ("_shift" if shift else "")] = {
            "ms": ms, "MSps_in": nvec * N / ms / 1e3, "alg_GBps": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / peak,
            "ms_for_1G_samples": ms * (2 ** 30) / (nvec * N)}
    # stand-alone discriminator on [time][channel] data
    rows, M = 12500, 8000
    xq = x[: (rows + 1) * M]
    d = torch.empty((rows, M), dtype=torch.float32, device=dev)
    q = B.quadrature_demod_cf(3.07)
    ms = timeit(lambda: q.work_device(rows, M, xq, d), 10, flush)
    out["quadrature_demod_cf_100M"] = {"ms": ms, "alg_GBps": 12.0 * rows * M / ms / 1e6, "frac_hbm": 12.0 * rows * M / ms / 1e6 / peak}
    del x, y, xq, d
    # pfb_arb_resampler_ccf batched over the cfg5 channel block: 12 500 rows x 8000 channels -> 19 200 rows
    # (12.5 kS/s -> 4 samples/symbol), 17 taps per filter; algorithmic bytes 8 B per input item + 8 B per output item
    rate = 19200.0 / 12500.0
    rtaps = np.asarray(firdes.low_pass(32, 32 * 12500.0, 5000.0, 2500.0), np.float32)
    arb = B.pfb_arb_resampler_ccf(rate, rtaps, 32, nchan=M)
    hrow = arb.history() - 1
    xin = torch.view_as_complex(torch.randn((hrow + rows, M, 2), generator=g, device=dev))
    yout = torch.empty((int(rows * rate) + 64, M), dtype=torch.complex64, device=dev)
    arb.work_device(1, 1, xin, yout)                       # the "updated" call
    produced = [0]

    def run_arb():
        n, _ = arb.work_device(yout.shape[0], xin.shape[0], xin, yout)
        produced[0] = n
    ms = timeit(run_arb, 10, flush)
    byt = 8.0 * rows * M + 8.0 * produced[0] * M
    out["pfb_arb_resampler_ccf_8000ch"] = {"ms": ms, "rows_out": produced[0], "taps_per_filter": arb.taps_per_filter(),
                                           "MSps_in": rows * M / ms / 1e3, "alg_GBps": byt / ms / 1e6,
                                           "frac_hbm": byt / ms / 1e6 / peak}
    del xin, yout
    # pfb_decimator_ccf: one channel out of M; algorithmic bytes 8 B per input sample (+ 8 B per output = 1/M of that)
    for Md, Td, tag in ((32, 16, "pfb_decimator_ccf_m32_t16"), (160, 16, "pfb_decimator_ccf_m160_t16")):
        nrow = 80_000_000 // Md
        dt = np.asarray(firdes.low_pass(1.0, float(Md), 0.4, 0.2), np.float32)
        dt = np.resize(dt, Md * Td).astype(np.float32)
        dec = B.pfb_decimator_ccf(Md, dt, 3)
        xr = torch.view_as_complex(torch.randn((nrow + Td - 1, Md, 2), generator=g, device=dev))
        yo = torch.empty(nrow, dtype=torch.complex64, device=dev)
        dec.work_device(1, xr, yo)
        ms = timeit(lambda: dec.work_device(nrow, xr, yo), 10, flush)
        byt = 8.0 * nrow * Md + 8.0 * nrow
        out[tag] = {"ms": ms, "MSps_in": nrow * Md / ms / 1e3, "alg_GBps": byt / ms / 1e6, "frac_hbm": byt / ms / 1e6 / peak}
        del xr, yo
    # cfg3: 2 MS/s -> 160 x 12.5 kHz channels (16 taps/branch) + batched DMR demod + sync search, 10 s of signal per launch
    from grb200 import chain as _chain
    M3, T3, rows3 = 160, 16, 125000
    t3 = firdes.low_pass_2(1.0, M3 * 12500.0, 5500.0, 1500.0, 60.0, firdes.WIN_BLACKMAN_hARRIS)
    c3 = len(t3) // 2
    t3 = (np.asarray(t3[c3 - M3 * T3 // 2: c3 - M3 * T3 // 2 + M3 * T3]) * M3).astype(np.float32)
    ch3 = _chain.DmrChain(_chain.DmrChainConfig(M3, t3, max_rows_per_block=rows3))
    x3 = torch.view_as_complex(torch.randn((ch3.history_rows() + rows3, M3, 2), generator=g, device=dev))
    s3 = torch.cuda.current_stream().cuda_stream

    def run3():
        ch3.process_device(x3, rows3, s3)
        ch3.join(s3)
    ms = timeit(run3, 5, flush)
    out["cfg3_pfb160_dmr_chain_20M"] = {"ms": ms, "MSps_in": rows3 * M3 / ms / 1e3, "seconds_of_signal_per_second": 10.0 / (ms * 1e-3),
                                        "alg_GBps": 49.9 * rows3 * M3 / ms / 1e6, "frac_hbm": 49.9 * rows3 * M3 / ms / 1e6 / peak}
    del x3, ch3
    # ---- blocks of SURVEY 8f rank 4 + the stand-alone a12 / a14 forms ---------------------------------------------------
    # fft_filter_ccc: 1025 and 4096 complex taps on 16 M samples (fused overlap-save), 20 000 taps (FFT engine)
    for nt, nblk_target, tag in ((129, 16_000_000, "fft_filter_ccc_129taps"), (1025, 16_000_000, "fft_filter_ccc_1025taps"),
                                 (4096, 16_000_000, "fft_filter_ccc_4096taps"), (20000, 16_000_000, "fft_filter_ccc_20000taps")):
        rngt = np.random.default_rng(nt)
        ct = ((rngt.standard_normal(nt) + 1j * rngt.standard_normal(nt)) / np.sqrt(nt)).astype(np.complex64)
        ff = B.fft_filter_ccc(1, ct)
        ns = ff.output_multiple()
        n = nblk_target // ns * ns
        xin = torch.view_as_complex(torch.randn((n, 2), generator=g, device=dev))
        yo = torch.empty(n, dtype=torch.complex64, device=dev)
        ms = timeit(lambda: ff.work_device(n, xin, yo), 5, flush)
        out[tag] = {"ms": ms, "path": ff.path(), "MSps_in": n / ms / 1e3, "alg_GBps": 16.0 * n / ms / 1e6, "frac_hbm": 16.0 * n / ms / 1e6 / peak,
                    "direct_form_GMACs_equiv": n * nt / ms / 1e6}
        del xin, yo
    # clock_recovery_mm_cc batched: 8000 channels x 12 500 rows (a cfg5 block of complex baseband), 2.604 samples/symbol
    rows, M = 12500, 8000
    xin = torch.view_as_complex(torch.randn((rows, M, 2), generator=g, device=dev))
    mm = B.clock_recovery_mm_cc(2.6041667, 0.25 * 0.175 * 0.175, 0.5, 0.175, 0.005, nchan=M)
    mo = torch.empty((6000, M), dtype=torch.complex64, device=dev)
    cnt = torch.zeros(M, dtype=torch.int32, device=dev)
    # simple timing: one launch over the whole block from a fresh state each time
    def run_mmcc_fresh():
        blk = run_mmcc_fresh.blk
        blk.work_device(rows, run_mmcc_fresh.base, xin, mo, None, 6000, cnt)
        run_mmcc_fresh.base += rows - 24
    run_mmcc_fresh.blk = mm
    run_mmcc_fresh.base = 0
    ms = timeit(run_mmcc_fresh, 5, flush)
    nsym = float(cnt.float().mean().item())
    out["clock_recovery_mm_cc_8000ch"] = {"ms": ms, "symbols_per_channel": nsym, "MSps_in": rows * M / ms / 1e3,
                                          "alg_GBps": (8.0 * rows * M + 8.0 * nsym * M) / ms / 1e6,
                                          "frac_hbm": (8.0 * rows * M + 8.0 * nsym * M) / ms / 1e6 / peak,
                                          "cycles_per_symbol_at_1965MHz": ms * 1e-3 * 1965e6 / max(nsym, 1)}
    del xin, mo
    # framer_sink_1 batched behind the correlator's bytes: 8000 channels x 9600 bits, a flag every ~600 bits
    nbits = 9600
    by = (torch.rand((nbits, M), generator=g, device=dev) < 0.5).to(torch.uint8)
    by |= ((torch.rand((nbits, M), generator=g, device=dev) < 1.0 / 600).to(torch.uint8) << 1)
    fr = B.framer_sink_1(M, max_msgs=1 << 18, payload_capacity=1 << 26)
    ms = timeit(lambda: fr.work_device(nbits, by, M, 1), 5, flush)
    nmsg = len(fr.messages())
    out["framer_sink_1_8000ch"] = {"ms": ms, "bits": nbits * M, "messages_total": nmsg, "alg_GBps": 1.0 * nbits * M / ms / 1e6,
                                   "frac_hbm": 1.0 * nbits * M / ms / 1e6 / peak}
    # map_bb / unpack_k_bits_bb / stream_to_streams on 256 M items
    nb = 1 << 28
    xb = torch.randint(0, 4, (nb,), dtype=torch.uint8, device=dev)
    yb = torch.empty(nb, dtype=torch.uint8, device=dev)
    mp = B.map_bb([0, 1, 3, 2])
    ms = timeit(lambda: mp.work_device(nb, xb, yb), 5, flush)
    out["map_bb_256M"] = {"ms": ms, "alg_GBps": 2.0 * nb / ms / 1e6, "frac_hbm": 2.0 * nb / ms / 1e6 / peak}
    up = B.unpack_k_bits_bb(2)
    ms = timeit(lambda: up.work_device(nb, xb, yb), 5, flush)
    out["unpack_k_bits_bb_k2_256M_out"] = {"ms": ms, "alg_GBps": 1.5 * nb / ms / 1e6, "frac_hbm": 1.5 * nb / ms / 1e6 / peak}
    del xb, yb
    rows, M = 12500, 8000
    xin = torch.view_as_complex(torch.randn((rows, M, 2), generator=g, device=dev))
    yo = torch.empty((M, rows), dtype=torch.complex64, device=dev)
    st = B.stream_to_streams(8, M)
    ms = timeit(lambda: st.work_device(rows, xin, yo), 5, flush)
    out["stream_to_streams_8000x12500_c64"] = {"ms": ms, "alg_GBps": 16.0 * rows * M / ms / 1e6, "frac_hbm": 16.0 * rows * M / ms / 1e6 / peak}
    del xin, yo
    # raw copy rates of the box (pinned), 800 MB like one bench block
    h = torch.empty(100_000_000, dtype=torch.complex64, pin_memory=True)
    dd = torch.empty(100_000_000, dtype=torch.complex64, device=dev)
    for name, fn in (("h2d", lambda: dd.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(dd, non_blocking=True))):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        out["pinned_%s_GBps" % name] = 5 * 0.8 / (time.perf_counter() - t0)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
