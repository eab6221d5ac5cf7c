"""The C++ host blocks (include/gr_b200_blocks.h: the reference's gr_block interface over libgr_cuda) driven
like the reference's QA code drives its blocks -- vector_source -> block -> vector_sink under a small
single-threaded scheduler (tests/cpp/block_harness.cc) -- and compared with the oracle.

Not-GPU part: the header compiles with plain g++, links against libgr_cuda.so, and constructor argument
errors surface as the reference's exception types without touching a device."""
import os
import subprocess

import numpy as np
import pytest

from conftest import has_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200")
EXE = os.path.join(ROOT, "build", "block_harness")
TOL = 1e-4   # BASELINE.json north_star: max relative error on FIR / channelizer / FFT outputs


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "cpp", "block_harness.cc")
    deps = [src, os.path.join(ROOT, "include", "gr_b200_blocks.h"), os.path.join(ROOT, "include", "gr_b200_runtime.h"),
            os.path.join(ROOT, "include", "gr_cuda.h")]
    if not os.path.exists(os.path.join(PKG, "libgr_cuda.so")):
        pytest.skip("libgr_cuda.so not built")
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), src,
                               "-L" + PKG, "-lgr_cuda", "-Wl,-rpath," + PKG, "-o", EXE])
    return EXE


def relerr(a, b):
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def test_header_compiles_and_maps_argument_errors(harness):
    out = subprocess.check_output([harness, "errors"], text=True)
    got = dict(l.split() for l in out.strip().splitlines())
    assert got == {
        "pfb_bad_oversample": "invalid_argument",   # gr_pfb_channelizer_ccf.cc:57-60
        "mm_omega_lt_1": "out_of_range",            # digital_clock_recovery_mm_ff.cc:58-59
        "mm_negative_gain": "out_of_range",         # :60-61
        "corr_code_too_long": "out_of_range",       # digital_correlate_access_code_bb.cc:54-57
        "fft_size_zero": "out_of_range",            # gri_fft.cc:104-105
        "io_signature": "invalid_argument",
        "unpack_k_zero": "out_of_range",            # gr_unpack_k_bits_bb.cc:44-45
        "mmcc_omega_zero": "out_of_range",          # digital_clock_recovery_mm_cc.cc:65-66
        "mmcc_negative_gain": "out_of_range",       # :67-68
    }


gpu = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]


def run(harness, tmp_path, spec, x, out_dtype, max_noutput=1000, files=()):
    inp, outp = tmp_path / "in.bin", tmp_path / "out.bin"
    np.ascontiguousarray(x).tofile(inp)
    args = []
    for s in spec:
        if isinstance(s, np.ndarray):
            f = tmp_path / ("arg%d.bin" % len(args))
            np.ascontiguousarray(s, np.float32).tofile(f)
            args.append(str(f))
        else:
            args.append(str(s))
    subprocess.check_call([harness, "run"] + args + [str(inp), str(outp), str(max_noutput)])
    return np.fromfile(outp, dtype=out_dtype)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_scheduler_visible_contracts(harness):
    out = subprocess.run([harness, "contract"], text=True, capture_output=True)
    assert out.returncode == 0 and "contract ok" in out.stdout, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_fir_blocks_through_the_scheduler(harness, tmp_path, orc):
    rng = np.random.default_rng(21)
    x = (rng.standard_normal(20000) + 1j * rng.standard_normal(20000)).astype(np.complex64)
    taps = rng.standard_normal(64).astype(np.float32)
    y = run(harness, tmp_path, ["fir_ccf", 4, taps], x, np.complex64, max_noutput=777)
    want = orc.fir_ccf(taps, 4, x)
    assert len(y) >= len(want) - 1 and relerr(y, want[:len(y)]) < TOL
    xf = rng.standard_normal(30000).astype(np.float32)
    tf = rng.standard_normal(29).astype(np.float32)
    for chunk in (1000, 333):    # the SSE summation order follows the ABSOLUTE index: chunking must not matter
        yf = run(harness, tmp_path, ["fir_fff", 1, tf], xf, np.float32, max_noutput=chunk)
        wf = orc.fir_fff(tf, 1, xf, order=orc.ORDER_SSE)
        assert len(yf) == len(wf) and np.array_equal(yf, wf)
    yx = run(harness, tmp_path, ["fxlat", 5, taps, 12500.0, 100000.0], x, np.complex64, max_noutput=512)
    wx = orc.freq_xlating_fir_ccf(taps, 5, 12500.0, 100000.0, x)
    wx = wx[0] if isinstance(wx, tuple) else wx
    assert len(yx) >= len(wx) - 1 and relerr(yx, wx[:len(yx)]) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("M,T,os_rate", [(20, 6, 1.0), (160, 16, 1.0), (8, 4, 2.0)])
def test_pfb_channelizer_block(harness, tmp_path, orc, M, T, os_rate):
    rng = np.random.default_rng(M)
    rows = 600
    x = (rng.standard_normal(rows * M) + 1j * rng.standard_normal(rows * M)).astype(np.complex64)
    taps = rng.standard_normal(M * T - 3).astype(np.float32)
    y = run(harness, tmp_path, ["pfb", M, taps, os_rate], x, np.complex64, max_noutput=96)
    want, _ = orc.pfb_channelizer_ccf(M, taps, x, os_rate)
    want = want.reshape(-1)
    n = min(len(y), len(want))
    assert n >= len(want) - 96 * M and n > 0
    assert relerr(y[:n], want[:n]) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_pfb_decimator_block(harness, tmp_path, orc):
    """gr_make_pfb_decimator_ccf under the scheduler: decim input streams (gr_stream_to_streams), a sync block with
    history = taps_per_filter."""
    rng = np.random.default_rng(29)
    for M, T, ch in ((8, 5, 3), (160, 16, 7)):
        x = (rng.standard_normal(M * 900) + 1j * rng.standard_normal(M * 900)).astype(np.complex64)
        taps = (rng.standard_normal(M * T - 2) * 0.1).astype(np.float32)
        y = run(harness, tmp_path, ["pfbdec", M, taps, ch], x, np.complex64, max_noutput=256)
        want = orc.pfb_decimator_ccf(M, taps, ch, x)
        assert len(want) - 256 <= len(y) <= len(want) and relerr(y, want[:len(y)]) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_fft_filter_block(harness, tmp_path, orc):
    """gr_make_fft_filter_ccc under the scheduler: history 1, output_multiple = nsamples, decimation 2."""
    rng = np.random.default_rng(31)
    x = (rng.standard_normal(20000) + 1j * rng.standard_normal(20000)).astype(np.complex64)
    taps = ((rng.standard_normal(45) + 1j * rng.standard_normal(45)) * 0.1).astype(np.complex64)
    y = run(harness, tmp_path, ["fftfilt", 2, taps.view(np.float32)], x, np.complex64, max_noutput=1000)
    want = orc.FftFilter(2, taps).run(x)
    assert len(want) - 2000 <= len(y) <= len(want) and len(y) % 84 == 0     # nsamples = 128 - 45 + 1
    assert relerr(y, want[:len(y)]) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_pfb_arb_resampler_block(harness, tmp_path, orc):
    """gr_make_pfb_arb_resampler_ccf under the scheduler: forecast = noutput + history - 1 (gr_block default),
    consume_each honoured, first general_work returns 0; bit identical to the generic-order reference."""
    rng = np.random.default_rng(23)
    x = (rng.standard_normal(12000) + 1j * rng.standard_normal(12000)).astype(np.complex64)
    taps = (rng.standard_normal(32 * 9 + 1) * 0.1).astype(np.float32)
    for rate, chunk in ((1.536, 600), (0.41, 333)):
        y = run(harness, tmp_path, ["arb", rate, taps, 32], x, np.complex64, max_noutput=chunk)
        want = orc.ArbResampler(rate, taps, 32).run(x)
        # the scheduler stops when forecast() no longer fits: the last few outputs are not asked for
        assert len(want) - 2 * chunk <= len(y) <= len(want)
        assert np.array_equal(y.view(np.uint32), want[:len(y)].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_fft_vcc_block(harness, tmp_path, orc):
    rng = np.random.default_rng(5)
    N, nvec = 4096, 40
    x = (rng.standard_normal(N * nvec) + 1j * rng.standard_normal(N * nvec)).astype(np.complex64)
    from grb200 import firdes
    w = np.asarray(firdes.window(firdes.WIN_BLACKMAN_hARRIS, N), np.float32)
    y = run(harness, tmp_path, ["fft", N, 1, w, 0], x, np.complex64, max_noutput=7)
    assert relerr(y, orc.fft_vcc(N, True, w, False, x)) < TOL
    y = run(harness, tmp_path, ["fft", N, 1, "-", 1], x, np.complex64, max_noutput=16)
    assert relerr(y, orc.fft_vcc(N, True, None, True, x)) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_demod_tail_blocks_bit_exact(harness, tmp_path, orc):
    """BASELINE config 2 shape: quadrature_demod_cf -> RRC -> clock_recovery_mm_ff -> 4-level slicer ->
    correlate_access_code_bb as separate C++ blocks, each bit exact against the oracle on the oracle's input."""
    from grb200 import firdes, synth
    rng = np.random.default_rng(3)
    fs, sps = 48000.0, 10
    sym = rng.choice([-3.0, -1.0, 1.0, 3.0], 3000)
    up = np.repeat(sym, sps).astype(np.float32)
    ph = np.cumsum(2 * np.pi * 648.0 * up / fs)
    xc = np.exp(1j * ph).astype(np.complex64) + 0.01 * (rng.standard_normal(len(ph)) + 1j * rng.standard_normal(len(ph))).astype(np.complex64)
    gain = fs / (2 * np.pi * 648.0)
    d = run(harness, tmp_path, ["quad", repr(float(np.float32(gain)))], xc, np.float32, max_noutput=4096)
    wd = orc.quadrature_demod_cf(np.float32(gain), xc)
    assert np.array_equal(d, wd)
    rrc = np.asarray(firdes.root_raised_cosine(1.0, fs, 4800.0, 0.2, 11 * sps + 1), np.float32)
    f = run(harness, tmp_path, ["fir_fff", 1, rrc], wd, np.float32, max_noutput=3000)
    wf = orc.fir_fff(rrc, 1, wd, order=orc.ORDER_SSE)
    assert np.array_equal(f, wf)
    omega, gmu = float(sps), 0.175
    gom = 0.25 * gmu * gmu
    m = run(harness, tmp_path, ["mm", omega, repr(gom), 0.5, gmu, 0.005], wf, np.float32, max_noutput=500)
    wm, _ = orc.mm_work(orc.mm_new(omega, gom, 0.5, gmu, 0.005), wf, order=orc.ORDER_SSE)
    assert len(m) > 2900 and np.array_equal(m, wm[:len(m)]) and len(m) >= len(wm) - 2
    s = run(harness, tmp_path, ["slicer4", 0.0], wm, np.uint8, max_noutput=999)
    ws = orc.slicer4(wm, 0.0)
    assert np.array_equal(s, ws)
    s2 = run(harness, tmp_path, ["slicer2"], wm, np.uint8, max_noutput=999)
    assert np.array_equal(s2, orc.binary_slicer(wm))
    bits = orc.unpack_k_bits_bb(2, orc.map_bb(synth.SLICER_TO_DIBIT_MAP, ws))
    code = synth.access_code_string(synth.DMR_BS_DATA_SYNC_BITS)
    bits[1000:1048] = np.frombuffer(code.encode(), np.uint8) & 1      # plant one sync word
    c = run(harness, tmp_path, ["corr", code, 1], bits, np.uint8, max_noutput=640)
    wc = orc.corr_work(orc.corr_new(code, 1), bits)
    assert np.array_equal(c, wc) and np.any(c & 2)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_next_blocks_through_the_cpp_mirror(harness, tmp_path, orc, golden_next):
    """gr_map_bb, gr_unpack_k_bits_bb, gr_stream_to_streams / gr_vector_to_streams, digital_clock_recovery_mm_cc and
    gr_framer_sink_1 (into a gr_msg_queue) as C++ blocks under the scheduler loop, against the oracle / fixtures."""
    rng = np.random.default_rng(12)
    x = rng.integers(0, 4, 5001).astype(np.uint8)
    m = np.array([0, 1, 3, 2], np.float32)
    y = run(harness, tmp_path, ["map", m], x, np.uint8, max_noutput=777)
    assert np.array_equal(y, orc.map_bb([0, 1, 3, 2], x))
    u = run(harness, tmp_path, ["unpack", 2], y, np.uint8, max_noutput=1001)       # the scheduler rounds to multiples of k
    assert np.array_equal(u, orc.unpack_k_bits_bb(2, y))
    z = (rng.standard_normal(160 * 50) + 1j * rng.standard_normal(160 * 50)).astype(np.complex64)
    for kind in ("s2s", "v2s"):
        o = run(harness, tmp_path, [kind, 8, 160], z, np.complex64, max_noutput=17)
        assert np.array_equal(o.reshape(160, 50), z.reshape(50, 160).T)
    fx = golden_next
    args = [repr(float(v)) for v in fx["mmcc_args"]]
    s = run(harness, tmp_path, ["mmcc"] + args, fx["mmcc_x"], np.complex64, max_noutput=300)
    # history 3: the runtime presents two zero items in front of the stream (gr_buffer.cc:201-214)
    st = orc.mmcc_new(*[float(v) for v in fx["mmcc_args"]])
    want, _, _ = orc.mmcc_work(st, np.concatenate([np.zeros(2, np.complex64), fx["mmcc_x"]]))
    assert len(s) >= len(want) - 8 and np.array_equal(s.view(np.uint32), want[:len(s)].view(np.uint32))
    rec = run(harness, tmp_path, ["framer"], fx["framer_stream"], np.uint8, max_noutput=1500)
    got, pos = [], 0
    while pos < len(rec):
        n = int(rec[pos + 1]) | (int(rec[pos + 2]) << 8)
        got.append((int(rec[pos]), bytes(rec[pos + 3: pos + 3 + n])))
        pos += 3 + n
    want, p = [], 0
    for off, n in zip(fx["framer_offsets"], fx["framer_lengths"]):
        want.append((int(off), bytes(fx["framer_payloads"][p:p + n])))
        p += n
    assert got == want


def test_header_compiles_against_the_reference_runtime(tmp_path):
    """-DGR_B200_USE_GNURADIO_RUNTIME: every block derives from the reference's REAL gr_block / gr_sync_block /
    gr_sync_decimator / gr_sync_interpolator and uses its real gr_msg_queue / gr_message, compiled from the headers where
    they lie under /root/reference (Boost and UHD names come from tests/cpp/gr_shim: they are absent from this image).
    A signature that did not match the base class's virtual would hide it: -Werror=overloaded-virtual."""
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("/root/reference is not mounted on this box")
    core = os.path.join(ref, "gnuradio-core", "src", "lib")
    src = tmp_path / "real_runtime.cc"
    src.write_text("""#define GR_B200_USE_GNURADIO_RUNTIME
#include "gr_b200_blocks.h"
using namespace gr_b200;
// every factory is odr-used, so every constructor and member function body is compiled
void* factories[] = {(void*)&gr_make_fir_filter_ccf, (void*)&gr_make_fir_filter_fff, (void*)&gr_make_freq_xlating_fir_filter_ccf,
  (void*)&gr_make_pfb_channelizer_ccf, (void*)&gr_make_fft_filter_ccc, (void*)&gr_make_pfb_decimator_ccf,
  (void*)&gr_make_pfb_arb_resampler_ccf, (void*)&gr_make_fft_vcc, (void*)&gr_make_quadrature_demod_cf,
  (void*)&digital_make_clock_recovery_mm_ff, (void*)&pager_make_slicer_fb, (void*)&digital_make_binary_slicer_fb,
  (void*)&digital_make_correlate_access_code_bb, (void*)&digital_make_clock_recovery_mm_cc, (void*)&gr_make_framer_sink_1,
  (void*)&gr_make_map_bb, (void*)&gr_make_unpack_k_bits_bb, (void*)&gr_make_stream_to_streams, (void*)&gr_make_vector_to_streams};
gr_block* as_block(gr_pfb_channelizer_ccf* b) { return b; }   // really a gr_block of the reference
gr_sync_block* as_sync(gr_framer_sink_1* b) { return b; }
gr_sync_interpolator* as_interp(gr_unpack_k_bits_bb* b) { return b; }
int main() { return 0; }
""")
    cmd = ["g++", "-std=c++17", "-c", "-Wall", "-Werror=overloaded-virtual", "-w", "-Werror=overloaded-virtual",
           "-I" + os.path.join(ROOT, "tests", "cpp", "gr_shim"), "-I" + os.path.join(ROOT, "include"),
           "-I" + os.path.join(core, "runtime"), "-I" + os.path.join(core, "general"),
           "-I" + os.path.join(ref, "gruel", "src", "include"), str(src), "-o", str(tmp_path / "real_runtime.o")]
    r = subprocess.run(cmd, text=True, capture_output=True)
    assert r.returncode == 0, r.stderr[-3000:]
