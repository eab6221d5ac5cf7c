#!/usr/bin/env python3
"""Generates the committed fixtures in tests/golden/ (run HERE, where /root/reference is mounted).

  kats.json          known-answer vectors lifted from the reference's OWN tests for this path
                     (parsed out of the mounted test sources; file:line recorded per entry).
  ref_fixtures.npz   outputs of the reference's own classes (oracle/_ref/libgrref.so, i.e. the
                     unmodified sources compiled in place) on small seeded inputs, for the rows
                     no reference test pins (PFB channelizer, freq-xlating FIR, quadrature demod /
                     fast_atan2f, 4-level slicer, full demod chain) -- SURVEY.md 8c.

Neither file is read by the product; tests/ compare the oracle restatement AND the CUDA path to them.
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("GR_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))


def rd(rel):
    return open(os.path.join(REF, rel)).read()


def c_array(src, name):
    m = re.search(name + r"\s*\[[^\]]*\]\s*=\s*\{([^}]*)\}", src)
    return [float(v) for v in re.findall(r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?", m.group(1))]


def kats():
    k = {}
    # --- qa_gr_fir_fff.cc:58-76 ------------------------------------------------------------
    s = rd("gnuradio-core/src/lib/filter/qa_gr_fir_fff.cc")
    k["fir_fff"] = {
        "source": "gnuradio-core/src/lib/filter/qa_gr_fir_fff.cc:58-76",
        "input_1": c_array(s, "input_1"), "taps_1a": c_array(s, "taps_1a"), "expected_1a": c_array(s, "expected_1a"),
        "taps_1b": c_array(s, "taps_1b"), "expected_1b": c_array(s, "expected_1b"),
    }
    # --- qa_fft.py:26-33,50-100 -------------------------------------------------------------
    s = rd("gnuradio-core/src/python/gnuradio/gr/qa_fft.py")
    primes = [int(v) for v in re.findall(r"\d+", s[s.index("primes = ("):s.index(")", s.index("primes = ("))])]
    t1 = s[s.index("def test_001"):s.index("def test_002")]
    exp = t1[t1.index("expected_result = ("):t1.index("src = gr.vector_source_c")]
    vals = re.findall(r"\(([-+]?[\d.]+)([-+][\d.]+)j\)", exp)
    k["fft_vcc_32"] = {
        "source": "gnuradio-core/src/python/gnuradio/gr/qa_fft.py:50-100 (forward) and :101-158 (inverse)",
        "rel_eps": 4e-4, "abs_eps": 1e-9, "primes": primes[:64],
        "expected_re": [float(a) for a, b in vals], "expected_im": [float(b) for a, b in vals],
    }
    assert len(vals) == 32
    # --- qa_gr_firdes.cc t1 / t4 -------------------------------------------------------------
    s = rd("gnuradio-core/src/lib/general/qa_gr_firdes.cc")
    k["firdes_low_pass"] = {"source": "gnuradio-core/src/lib/general/qa_gr_firdes.cc:57-112,486-503",
                            "args": [1.0, 8000, 1750, 500], "win": 0, "expected": c_array(s, "t1_exp")}
    k["firdes_low_pass_2"] = {"source": "gnuradio-core/src/lib/general/qa_gr_firdes.cc:549-566",
                              "args": [1.0, 8000, 1750, 500, 66], "win": 0, "expected": c_array(s, "t4_exp")}
    # --- qa_correlate_access_code.py:50-78 ----------------------------------------------------
    k["correlate_access_code"] = {
        "source": "gr-digital/python/qa_correlate_access_code.py:27,50-78",
        "t1_code": "1011", "t1_src": [1, 0, 1, 1, 1, 1, 0, 1, 1] + [0] * 64 + [0] * 7,
        "t1_expected": [0] * 64 + [1, 0, 1, 1, 3, 1, 0, 1, 1, 2] + [0] * 6,
        "default_access_code_bytes": [0xAC, 0xDD, 0xA4, 0xE2, 0xF2, 0x8C, 0x20, 0xFC],
    }
    # --- qa_clock_recovery_mm.py:70-102,140-172 ------------------------------------------------
    k["clock_recovery_mm_ff"] = {
        "source": "gr-digital/python/qa_clock_recovery_mm.py:70-102,140-172",
        "test02": {"args": [2, 0.01, 0.5, 0.01, 0.001], "input": "100 x 1.0", "expected_last30": 0.99972,
                   "places": 5},
        "test04": {"args": [2, 0.01, 0.25, 0.1, 0.001], "input": "1000 x [1,1,-1,-1]", "expected_pm": 1.31,
                   "places": 1},
    }
    # --- qa_gr_math.cc:27-50 -----------------------------------------------------------------
    k["binary_slicer"] = {"source": "gnuradio-core/src/lib/general/qa_gr_math.cc:27-50",
                          "x": [-1, -0.5, 0, 0.5, 1.0], "z": [0, 0, 1, 1, 1]}
    k["binary_slicer_fb"] = {"source": "gr-digital/python/qa_binary_slicer_fb.py:35-50",
                             "src_sign": [-1, 1, -1, -1, 1, 1, -1, -1, -1, 1, 1, 1, -1, 1, 1, 1, 1],
                             "expected": [0, 1, 0, 0, 1, 1, 0, 0, 0, 1, 1, 1, 0, 1, 1, 1, 1]}
    k["rotator"] = {"source": "gnuradio-core/src/lib/filter/qa_gr_rotator.cc:43-75", "N": 100000,
                    "phase_incr": "2*pi/1003", "tol": 1e-4}
    k["mmse_interpolator"] = {"source": "gnuradio-core/src/lib/filter/qa_gri_mmse_fir_interpolator.cc:37-61",
                              "tol": 0.004}
    return k


def fixtures():
    import refharness as R
    from grb200 import synth
    fx = {}
    rng = np.random.default_rng(20261018)

    def crandn(n):
        return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)

    # a1: fir_filter_ccf, 64 taps, decim 4 (cfg1 shape, small), SSE and generic witnesses
    taps = rng.uniform(-1, 1, 64).astype(np.float32)
    x = crandn(4096)
    fx["fir_ccf_taps"], fx["fir_ccf_x"] = taps, x
    for impl, nm in ((1, "sse"), (0, "generic")):
        R.set_fir_impl(impl)
        fx["fir_ccf_y_" + nm] = R.run_sync(R.fir_filter_ccf(4, taps), x, decim=4)
    # a2: fir_filter_fff RRC
    rrc = R.firdes_root_raised_cosine(1.0, 12500.0, 4800.0, 0.2, 29)
    xf = rng.standard_normal(3000).astype(np.float32)
    fx["fir_fff_taps"], fx["fir_fff_x"] = rrc, xf
    for impl, nm in ((1, "sse"), (0, "generic")):
        R.set_fir_impl(impl)
        fx["fir_fff_y_" + nm] = R.run_sync(R.fir_filter_fff(1, rrc), xf)
    R.set_fir_impl(1)
    # a3: freq_xlating_fir_filter_ccf
    proto = R.firdes_low_pass(1.0, 2e6, 6000.0, 4000.0)[:257]
    xx = crandn(8192)
    fx["fx_proto"], fx["fx_x"] = proto, xx
    fx["fx_args"] = np.array([16, 250e3, 2e6])
    fx["fx_y"] = R.run_sync(R.freq_xlating_fir_filter_ccf(16, proto, 250e3, 2e6), xx, decim=16)
    # a4: pfb channelizer: M=160 T=16 (cfg3 shape), M=20 ragged taps, M=8 oversampled x2
    for tag, M, ntaps, osr, rows in (("m160", 160, 2560, 1.0, 40), ("m20", 20, 173, 1.0, 64), ("m8os2", 8, 32, 2.0, 48),
                                     ("m10os5", 10, 57, 5.0, 30)):
        t = (rng.standard_normal(ntaps) / ntaps).astype(np.float32)
        xin = crandn(M * rows)
        y, c = R.run_pfb(R.pfb_channelizer_ccf(M, t, osr), xin, M)
        fx["pfb_%s_taps" % tag], fx["pfb_%s_x" % tag], fx["pfb_%s_y" % tag] = t, xin, y
        fx["pfb_%s_meta" % tag] = np.array([M, osr, c])
    # a6: fft_vcc 4096 Blackman-Harris (window from gr_firdes::window, recorded choice), fwd shift on/off
    w = R.firdes_window(R.WIN_BLACKMAN_HARRIS, 4096)
    xv = crandn(4096 * 2)
    fx["fft4096_win"], fx["fft4096_x"] = w, xv
    fx["fft4096_y"] = R.run_sync(R.fft_vcc(4096, True, w, False), xv, vlen_in=4096)
    fx["fft4096_y_shift"] = R.run_sync(R.fft_vcc(4096, True, w, True), xv, vlen_in=4096)
    xs = crandn(160 * 3)
    fx["fft160_x"] = xs
    fx["fft160_y_inv_shift"] = R.run_sync(R.fft_vcc(160, False, [], True), xs, vlen_in=160)
    # a7/a8: quadrature demod + fast atan2 (incl. axes, zeros, tiny ratios, octant edges)
    yy = np.concatenate([rng.standard_normal(4000), [0, 0, 1, -1, 0, 0, 1e-3, 1, -1, 1, 3e-3, 1.0]]).astype(np.float32)
    xa = np.concatenate([rng.standard_normal(4000), [0, 1, 0, 0, -1, -0.0, 1, 1e-3, -1, -1, 1.0, 3.9e-3]]).astype(np.float32)
    fx["atan_y"], fx["atan_x"], fx["atan_out"] = yy, xa, R.fast_atan2f(yy, xa)
    xq = crandn(5000)
    fx["quad_x"], fx["quad_gain"] = xq, np.float32(12500.0 / (2 * np.pi * 648.0))
    fx["quad_y"] = R.run_sync(R.quadrature_demod_cf(float(fx["quad_gain"])), xq)
    # a9/a10: M&M on a noisy 4-level signal at 2.604 sps, SSE + generic
    sym = rng.integers(0, 4, 1500) * 2 - 3
    sig = synth.shape_symbols(sym, 12500.0 / 4800.0, rrc) + 0.05 * rng.standard_normal(int(1500 * 12500 / 4800) + 1)[: None]
    sig = sig.astype(np.float32)
    fx["mm_x"] = sig
    fx["mm_args"] = np.array([12500.0 / 4800.0, 0.25 * 0.175 * 0.175, 0.5, 0.175, 0.005], np.float32)
    for impl, nm in ((1, "sse"), (0, "generic")):
        R.set_fir_impl(impl)
        o, c = R.run_mm(R.clock_recovery_mm_ff(*[float(v) for v in fx["mm_args"]]), sig)
        fx["mm_y_" + nm], fx["mm_consumed_" + nm] = o, np.int64(c)
    R.set_fir_impl(1)
    # a11: 4-level slicer with and without DC tracking
    sl = (rng.standard_normal(2000) * 2.0).astype(np.float32)
    fx["slicer_x"] = sl
    fx["slicer_y_a0"] = R.run_sync(R.pager_slicer_fb(0.0), sl)
    fx["slicer_y_a01"] = R.run_sync(R.pager_slicer_fb(0.01), sl)
    # a13: correlator with a 48-bit DMR-style sync word and threshold 2
    code = synth.DMR_BS_DATA_SYNC_BITS
    bits = rng.integers(0, 2, 3000).astype(np.uint8)
    for pos in (100, 1000, 2500):
        bits[pos:pos + 48] = code
    bits[1010] ^= 1
    bits[2503] ^= 1; bits[2520] ^= 1; bits[2530] ^= 1
    fx["corr_bits"] = bits
    fx["corr_out_t2"] = R.run_sync(R.correlate_access_code_bb("".join(str(b) for b in code), 2), bits)
    return fx


def fixtures_next():
    """SURVEY.md 8f blocks (the callers either side of the path), same recipe: the reference's own classes on small
    seeded inputs.  Kept in a file of its own so that ref_fixtures.npz stays byte-identical."""
    import refharness as R
    from grb200 import firdes
    fx = {}
    rng = np.random.default_rng(20)
    # rank 3: gr_pfb_arb_resampler_ccf, the 12.5 kS/s -> 4 samples/symbol resampler of a 4FSK chain (rate 1.536), a
    # decimating rate, and ragged tap counts (taps_per_filter padding); SSE filters (what x86-64 runs) and generic
    x = (rng.standard_normal(3000) + 1j * rng.standard_normal(3000)).astype(np.complex64)
    fx["arb_x"] = x
    for tag, rate, nf, taps in (
            ("up", 19200.0 / 12500.0, 32, np.asarray(firdes.low_pass(32, 32 * 12500.0, 5000.0, 2500.0), np.float32)),
            ("down", 0.731, 32, (rng.standard_normal(32 * 9 + 5) * 0.1).astype(np.float32)),
            ("few", 2.5, 16, (rng.standard_normal(77) * 0.1).astype(np.float32))):
        fx["arb_%s_taps" % tag] = taps
        fx["arb_%s_args" % tag] = np.array([rate, nf])
        for impl, nm in ((1, "sse"), (0, "generic")):
            R.set_fir_impl(impl)
            fx["arb_%s_y_%s" % (tag, nm)] = R.run_arb(R.pfb_arb_resampler_ccf(rate, taps, nf), x, rate, chunk_out=211)
    R.set_fir_impl(1)
    # rank 3: gr_pfb_decimator_ccf, one channel out of 10 / 32 (tile path) and out of 160 (cfg3's channel count)
    for tag, M, T, ch in (("m10", 10, 7, 3), ("m32", 32, 12, 31), ("m160", 160, 16, 7)):
        taps = (rng.standard_normal(M * T - 3) * 0.1).astype(np.float32)
        xin = (rng.standard_normal(M * 120) + 1j * rng.standard_normal(M * 120)).astype(np.complex64)
        fx["pfbdec_%s_taps" % tag], fx["pfbdec_%s_x" % tag] = taps, xin
        fx["pfbdec_%s_args" % tag] = np.array([M, ch])
        fx["pfbdec_%s_y" % tag] = R.run_pfb_decimator(R.pfb_decimator_ccf(M, taps, ch), xin, M, chunk=50)
    # rank 4: gr_fft_filter_ccc (overlap-add, complex taps): decimation 1 and 3, tap counts on both sides of a power of two
    xf = (rng.standard_normal(6000) + 1j * rng.standard_normal(6000)).astype(np.complex64)
    fx["fftfilt_x"] = xf
    for tag, dec, nt in (("d1_t33", 1, 33), ("d3_t64", 3, 64), ("d1_t200", 1, 200)):
        tc = ((rng.standard_normal(nt) + 1j * rng.standard_normal(nt)) * 0.1).astype(np.complex64)
        blk = R.fft_filter_ccc(dec, tc)
        fx["fftfilt_%s_taps" % tag] = tc
        fx["fftfilt_%s_args" % tag] = np.array([dec, blk.output_multiple])
        fx["fftfilt_%s_y" % tag] = R.run_fft_filter(blk, xf, dec, blocks_per_call=2)
    # rank 4: gr_framer_sink_1 on a correlator-style stream: good packets (incl. zero length and the 4095-byte maximum),
    # every third header corrupted; packets recorded as [whitener offset, length, payload...] rows
    import orc
    pk = [(int(rng.integers(0, 16)), bytes(rng.integers(0, 256, int(n)).astype(np.uint8))) for n in (5, 0, 17, 1, 300, 4095, 2, 64)]
    stream = orc.framer_make_stream(rng, pk, corrupt_header_every=3)
    f = R.FramerSink()
    got = []
    for chunk in np.array_split(stream, 23):
        got += f.work(chunk)
    fx["framer_stream"] = stream
    fx["framer_offsets"] = np.array([g[0] for g in got], np.int32)
    fx["framer_lengths"] = np.array([len(g[1]) for g in got], np.int32)
    fx["framer_payloads"] = np.frombuffer(b"".join(g[1] for g in got), np.uint8)
    # rank 4: digital_clock_recovery_mm_cc on noisy QPSK at 4 samples/symbol, with and without the error output
    # (generic-order interpolator filters: the restatement's order)
    nq = 6000
    symq = (rng.integers(0, 2, nq // 4 + 1) * 2 - 1) + 1j * (rng.integers(0, 2, nq // 4 + 1) * 2 - 1)
    xq = (np.repeat(symq, 4)[:nq] + 0.1 * (rng.standard_normal(nq) + 1j * rng.standard_normal(nq))).astype(np.complex64)
    fx["mmcc_x"], fx["mmcc_args"] = xq, np.array([4.0, 0.25 * 0.1 * 0.1, 0.5, 0.1, 0.005])
    R.set_fir_impl(0)
    for we, nm in ((False, "plain"), (True, "err")):
        y, e, c = R.run_mm_cc(R.clock_recovery_mm_cc(*[float(v) for v in fx["mmcc_args"]]), xq, with_error=we)
        fx["mmcc_y_" + nm], fx["mmcc_consumed_" + nm] = y, np.int64(c)
        if we:
            fx["mmcc_err"] = e
    R.set_fir_impl(1)
    return fx


def fixtures_optfir():
    """Host-side design fixtures: taps of the reference's gr_remez (compiled in place) for the specs
    optfir.low_pass / high_pass / band_pass hand it, and the reference's own window.blackmanharris (the Python source of
    gnuradio/window.py:152-166 executed as it stands)."""
    import re
    sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))
    import refharness as R
    from grb200 import optfir
    out = {}
    specs = {"lp8": ("low", 1, 8, 0.4, 0.6, 0.1, 60), "lp32": ("low", 1, 32, 0.4, 0.6, 0.1, 100),
             "lp48k": ("low", 2.0, 48000.0, 3000.0, 4500.0, 0.2, 70), "hp": ("high", 1, 1.0, 0.2, 0.3, 0.1, 60)}
    for name, sp in specs.items():
        kind, gain, fs, f1, f2, rip, att = sp
        pd, sd = optfir.passband_ripple_to_dev(rip), optfir.stopband_atten_to_dev(att)
        if kind == "low":
            n, fo, ao, w = optfir.remezord([f1, f2], (gain, 0), [pd, sd], fs)
        else:
            n, fo, ao, w = optfir.remezord([f1, f2], (0, 1), [sd, pd], fs)
            if (n + 2) % 2 == 1:
                n += 1
        out["remez_" + name] = R.remez(n + 2, fo, ao, w)
        out["spec_" + name] = np.array([{"low": 0, "high": 1}[kind], gain, fs, f1, f2, rip, att], np.float64)
    src = rd("gnuradio-core/src/python/gnuradio/window.py")
    m = re.search(r"def coswindow\(coeffs\):.*?return closure\n", src, re.S)
    ns = {"math": __import__("math")}
    exec(m.group(0), ns)
    coeffs = re.search(r"blackmanharris = coswindow\(\((.*?)\)\)", src).group(1)
    bh = ns["coswindow"](tuple(float(v) for v in coeffs.split(",")))
    out["blackmanharris_64"] = np.array(bh(64), np.float64)
    out["blackmanharris_4096"] = np.array(bh(4096), np.float64)
    np.savez_compressed(os.path.join(HERE, "ref_fixtures_optfir.npz"), **out)
    print("wrote ref_fixtures_optfir.npz:", sorted(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "optfir":
        fixtures_optfir()
        return
    if len(sys.argv) > 1 and sys.argv[1] == "next":
        np.savez_compressed(os.path.join(HERE, "ref_fixtures_next.npz"), **fixtures_next())
        print("wrote ref_fixtures_next.npz", os.path.getsize(os.path.join(HERE, "ref_fixtures_next.npz")) // 1024, "KiB")
        return
    with open(os.path.join(HERE, "kats.json"), "w") as f:
        json.dump(kats(), f, indent=1)
    np.savez_compressed(os.path.join(HERE, "ref_fixtures.npz"), **fixtures())
    print("wrote kats.json, ref_fixtures.npz",
          os.path.getsize(os.path.join(HERE, "ref_fixtures.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
