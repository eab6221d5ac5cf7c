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
