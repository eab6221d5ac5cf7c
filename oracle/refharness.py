"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package (grb200).

ctypes harness over oracle/_ref/libgrref.so, i.e. the reference's own C++/asm classes
compiled from /root/reference by oracle/build_ref.sh.  It plays the role of the reference
runtime around a block: it lays out history-prefixed, alignment-controlled input buffers
exactly as gr_flat_flowgraph / gr_buffer would (gnuradio-core/src/lib/runtime/
gr_flat_flowgraph.cc:150, gr_buffer.cc:201-214: history-1 zero items pre-loaded, buffer base
page aligned so that absolute item index a sits at byte offset a*itemsize mod 16) and calls
general_work() once (or in chunks) on them.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "_ref", "libgrref.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/_ref/libgrref.so missing: run oracle/build_ref.sh (needs /root/reference)")
        L = C.CDLL(path)
        L.grref_last_error.restype = C.c_char_p
        for name in dir_makers():
            getattr(L, name).restype = C.c_void_p
        L.grref_block_relative_rate.restype = C.c_double
        L.grref_block_history.restype = C.c_uint
        L.grref_pager_slicer_dc_offset.restype = C.c_float
        L.grref_branchless_clip.restype = C.c_float
        L.grref_branchless_clip.argtypes = [C.c_float, C.c_float]
        L.grref_binary_slicer.argtypes = [C.c_float]
        L.grref_count_bits64.argtypes = [C.c_ulonglong]
        _LIB = L
    return _LIB


def dir_makers():
    return [
        "grref_make_fir_filter_ccf", "grref_make_fir_filter_fff", "grref_make_freq_xlating_fir_filter_ccf",
        "grref_make_pfb_channelizer_ccf", "grref_make_fft_vcc", "grref_make_quadrature_demod_cf",
        "grref_make_clock_recovery_mm_ff", "grref_make_pager_slicer_fb", "grref_make_binary_slicer_fb",
        "grref_make_map_bb", "grref_make_unpack_k_bits_bb", "grref_make_correlate_access_code_bb",
        "grref_make_pfb_arb_resampler_ccf", "grref_make_pfb_decimator_ccf", "grref_make_fft_filter_ccc", "grref_make_framer_sink_1", "grref_make_clock_recovery_mm_cc",
    ]


def available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libgrref.so"))


def set_fir_impl(sse):
    """1 -> the SSE classes gr_fir_sysconfig_x86 selects on x86-64; 0 -> *_generic."""
    lib().grref_set_fir_impl(int(bool(sse)))


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def aligned_stream(new_items, history, dtype):
    """Buffer = (history-1) zeros + new items, placed so that new item 0 is 16-byte aligned
    (absolute item index a -> address = a*itemsize mod 16, as in a page-aligned gr_buffer).
    Returns (backing array, pointer-to-first-history-item as int, view incl. history)."""
    new_items = np.ascontiguousarray(new_items, dtype=dtype)
    isz = new_items.dtype.itemsize
    h = history - 1
    pad_items = (16 // isz) * 4 + 16  # slack: SSE kernels read up to 15 B before the pointer
    total = pad_items + h + len(new_items) + pad_items
    raw = np.zeros(total * isz + 64, dtype=np.uint8)
    base = raw.ctypes.data
    # choose start so that address of new item 0 is 16-aligned
    off = pad_items * isz
    addr0 = base + off + h * isz
    off += (-addr0) % 16
    view = raw[off: off + (h + len(new_items)) * isz].view(dtype)
    view[h:] = new_items
    return raw, base + off, view


class RefBlock:
    """One reference block instance (kept alive across work calls, like in a flowgraph)."""

    def __init__(self, handle, in_dtype, out_dtype, nin=1, out_vlen=1):
        if not handle:
            raise_from_ref()
        self.h = C.c_void_p(handle)
        self.in_dtype = np.dtype(in_dtype)
        self.out_dtype = np.dtype(out_dtype)
        self.nin = nin
        self.out_vlen = out_vlen

    def __del__(self):
        try:
            lib().grref_block_delete(self.h)
        except Exception:
            pass

    history = property(lambda s: lib().grref_block_history(s.h))
    output_multiple = property(lambda s: lib().grref_block_output_multiple(s.h))
    relative_rate = property(lambda s: lib().grref_block_relative_rate(s.h))
    consumed = property(lambda s: lib().grref_block_consumed(s.h))

    def forecast(self, noutput):
        return lib().grref_block_forecast(self.h, noutput, self.nin)

    def general_work(self, noutput, in_ptrs, ninput_items, out):
        nin = len(in_ptrs)
        ip = (C.c_void_p * nin)(*in_ptrs)
        ni = (C.c_int * nin)(*ninput_items)
        op = (C.c_void_p * 1)(out.ctypes.data)
        return lib().grref_block_general_work(self.h, int(noutput), ni, nin, ip, op, 1)


class RefError(Exception):
    pass


def raise_from_ref():
    msg = lib().grref_last_error().decode()
    if msg.startswith("invalid_argument"):
        raise ValueError(msg)
    if msg.startswith("out_of_range"):
        raise IndexError(msg)
    raise RefError(msg)


# ---- block factories (reference names) -------------------------------------------------
def fir_filter_ccf(decim, taps):
    t = np.ascontiguousarray(taps, np.float32)
    return RefBlock(lib().grref_make_fir_filter_ccf(int(decim), _fp(t), len(t)), np.complex64, np.complex64)


def fir_filter_fff(decim, taps):
    t = np.ascontiguousarray(taps, np.float32)
    return RefBlock(lib().grref_make_fir_filter_fff(int(decim), _fp(t), len(t)), np.float32, np.float32)


def freq_xlating_fir_filter_ccf(decim, taps, center_freq, sampling_freq):
    t = np.ascontiguousarray(taps, np.float32)
    return RefBlock(lib().grref_make_freq_xlating_fir_filter_ccf(int(decim), _fp(t), len(t), C.c_double(center_freq),
                                                                 C.c_double(sampling_freq)), np.complex64, np.complex64)


def pfb_channelizer_ccf(numchans, taps, oversample_rate=1.0):
    t = np.ascontiguousarray(taps, np.float32)
    return RefBlock(lib().grref_make_pfb_channelizer_ccf(int(numchans), _fp(t), len(t), C.c_float(oversample_rate)),
                    np.complex64, np.complex64, nin=numchans, out_vlen=numchans)


def pfb_arb_resampler_ccf(rate, taps, filter_size=32):
    t = np.ascontiguousarray(taps, np.float32)
    return RefBlock(lib().grref_make_pfb_arb_resampler_ccf(C.c_float(rate), _fp(t), len(t), int(filter_size)),
                    np.complex64, np.complex64)


def run_arb(block, x, rate, chunk_out=None):
    """gr_pfb_arb_resampler_ccf over the whole stream x (new items): the scheduler's part is played here --
    history-prefixed aligned buffer, ninput_items counted from the first history item, consume_each honoured."""
    x = np.ascontiguousarray(x, np.complex64)
    hist = block.history
    raw, ptr, view = aligned_stream(x, hist, np.complex64)
    total = len(view)
    pos, outs, first = 0, [], True
    while True:
        avail = total - pos
        nout = chunk_out or int(avail * rate) + 16
        out = np.zeros(nout, np.complex64)
        r = block.general_work(nout, [ptr + pos * 8], [avail], out)
        c = block.consumed
        outs.append(out[:r].copy())
        pos += c
        if r == 0 and c == 0:
            if first:
                first = False
                continue
            break
        first = False
    return np.concatenate(outs) if outs else np.zeros(0, np.complex64)


def pfb_decimator_ccf(decim, taps, channel):
    t = np.ascontiguousarray(taps, np.float32)
    return RefBlock(lib().grref_make_pfb_decimator_ccf(int(decim), _fp(t), len(t), int(channel)), np.complex64, np.complex64,
                    nin=int(decim))


def run_pfb_decimator(block, x, decim, chunk=None):
    """gr_stream_to_streams -> gr_pfb_decimator_ccf over the interleaved stream x.  The first work() after set_taps
    returns 0 (gr_pfb_decimator_ccf.cc:136-139)."""
    x = np.ascontiguousarray(x, np.complex64)
    n = len(x) // decim
    hist = block.history
    keep = [aligned_stream(x[s::decim][:n], hist, np.complex64) for s in range(decim)]
    out = np.zeros(max(n, 1), np.complex64)
    assert block.general_work(min(n, 8), [k[1] for k in keep], [n + hist - 1] * decim, out) == 0
    done, step = 0, chunk or n
    while done < n:
        m = min(step, n - done)
        r = block.general_work(m, [k[1] + done * 8 for k in keep], [m + hist - 1] * decim, out[done:])
        assert r == m
        done += r
    return out[:n]


def fft_filter_ccc(decim, taps):
    t = np.ascontiguousarray(taps, np.complex64)
    return RefBlock(lib().grref_make_fft_filter_ccc(int(decim), _fp(t.view(np.float32)), len(t)), np.complex64, np.complex64)


def fft_filter_ccc_set_taps(block, taps):
    t = np.ascontiguousarray(taps, np.complex64)
    assert lib().grref_fft_filter_ccc_set_taps(block.h, _fp(t.view(np.float32)), len(t)) == 0


def run_fft_filter(block, x, decim, blocks_per_call=None):
    """gr_fft_filter_ccc over the new items x (history 1).  noutput is a multiple of output_multiple = nsamples
    (gr_fft_filter_ccc.cc:66,92-100); the tail that does not fill a whole block is not asked for, like the scheduler."""
    x = np.ascontiguousarray(x, np.complex64)
    ns = block.output_multiple
    raw, ptr, view = aligned_stream(x, 1, np.complex64)
    nblocks = (len(x) // decim) // ns
    out = np.zeros(max(nblocks * ns, 1), np.complex64)
    done, step = 0, (blocks_per_call or max(nblocks, 1)) * ns
    while done < nblocks * ns:
        n = min(step, nblocks * ns - done)
        r = block.general_work(n, [ptr + done * decim * 8], [n * decim], out[done:])
        assert r == n, (r, n)
        done += r
    return out[:nblocks * ns]


class FramerSink:
    """gr_framer_sink_1 + the gr_msg_queue it posts to."""

    def __init__(self):
        self.blk = RefBlock(lib().grref_make_framer_sink_1(), np.uint8, np.uint8)

    def __del__(self):
        try:
            lib().grref_framer_forget(self.blk.h)
        except Exception:
            pass

    def work(self, in_bytes):
        x = np.ascontiguousarray(in_bytes, np.uint8)
        if len(x):
            ip = (C.c_void_p * 1)(x.ctypes.data)
            ni = (C.c_int * 1)(len(x))
            r = lib().grref_block_general_work(self.blk.h, len(x), ni, 1, ip, None, 0)
            assert r == len(x), r
        got, buf, a1 = [], np.zeros(4096, np.uint8), C.c_double(0)
        while True:
            n = lib().grref_framer_pop(self.blk.h, buf.ctypes.data_as(C.c_void_p), 4096, C.byref(a1))
            if n < 0:
                break
            got.append((int(a1.value), bytes(buf[:n])))
        return got


def fft_vcc(fft_size, forward, window, shift=False):
    w = np.ascontiguousarray(window, np.float32)
    return RefBlock(lib().grref_make_fft_vcc(int(fft_size), int(forward), _fp(w), len(w), int(shift)),
                    np.complex64, np.complex64, out_vlen=fft_size)


def quadrature_demod_cf(gain):
    return RefBlock(lib().grref_make_quadrature_demod_cf(C.c_float(gain)), np.complex64, np.float32)


def clock_recovery_mm_ff(omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001):
    return RefBlock(lib().grref_make_clock_recovery_mm_ff(C.c_float(omega), C.c_float(gain_omega), C.c_float(mu),
                                                          C.c_float(gain_mu), C.c_float(omega_relative_limit)),
                    np.float32, np.float32)


def clock_recovery_mm_cc(omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001):
    return RefBlock(lib().grref_make_clock_recovery_mm_cc(C.c_float(omega), C.c_float(gain_omega), C.c_float(mu),
                                                          C.c_float(gain_mu), C.c_float(omega_relative_limit)),
                    np.complex64, np.complex64)


def run_mm_cc(block, x, noutput=None, with_error=False):
    """digital_clock_recovery_mm_cc.general_work on the complex stream x (one or two outputs).
    Returns (symbols, error signal or None, consumed)."""
    x = np.ascontiguousarray(x, np.complex64)
    raw, ptr, view = aligned_stream(x, 1, np.complex64)
    nout = noutput if noutput is not None else len(x)
    out = np.zeros(max(nout, 1), np.complex64)
    err = np.zeros(max(nout, 1), np.float32)
    ip = (C.c_void_p * 1)(ptr)
    ni = (C.c_int * 1)(len(x))
    if with_error:
        op = (C.c_void_p * 2)(out.ctypes.data, err.ctypes.data)
        r = lib().grref_block_general_work(block.h, int(nout), ni, 1, ip, op, 2)
    else:
        op = (C.c_void_p * 1)(out.ctypes.data)
        r = lib().grref_block_general_work(block.h, int(nout), ni, 1, ip, op, 1)
    return out[:r], (err[:r] if with_error else None), block.consumed


def pager_slicer_fb(alpha):
    return RefBlock(lib().grref_make_pager_slicer_fb(C.c_float(alpha)), np.float32, np.uint8)


def binary_slicer_fb():
    return RefBlock(lib().grref_make_binary_slicer_fb(), np.float32, np.uint8)


def map_bb(m):
    a = (C.c_int * len(m))(*[int(v) for v in m])
    return RefBlock(lib().grref_make_map_bb(a, len(m)), np.uint8, np.uint8)


def unpack_k_bits_bb(k):
    return RefBlock(lib().grref_make_unpack_k_bits_bb(int(k)), np.uint8, np.uint8)


def correlate_access_code_bb(code, threshold):
    return RefBlock(lib().grref_make_correlate_access_code_bb(code.encode(), int(threshold)), np.uint8, np.uint8)


# ---- "flowgraph" drivers: vector_source -> block -> vector_sink ------------------------
def run_sync(block, x, decim=1, vlen_in=1, chunk=None):
    """Run a gr_sync_block / gr_sync_decimator over the whole stream x (new items only).
    Returns all outputs.  `chunk` (in output items) emulates repeated work() calls; the
    block keeps its own state, the harness keeps the history like gr_buffer does."""
    hist = block.history
    item = block.in_dtype
    x = np.ascontiguousarray(x, item)
    n_items = len(x) // vlen_in
    raw, ptr, view = aligned_stream(x, (hist - 1) * vlen_in + 1, item)
    nout_total = n_items // decim
    out = np.zeros(nout_total * block.out_vlen, block.out_dtype)
    isz = item.itemsize * vlen_in
    done = 0
    step = chunk or nout_total
    while done < nout_total:
        n = min(step, nout_total - done)
        o = out[done * block.out_vlen:]
        r = block.general_work(n, [ptr + done * decim * isz], [n * decim + hist - 1], o)
        if r == 0 and n > 0:
            # "return 0 once after set_taps" contract: history may have changed; re-layout
            if block.history != hist:
                hist = block.history
                raw, ptr, view = aligned_stream(x, (hist - 1) * vlen_in + 1, item)
            continue
        done += r
    return out


def run_mm(block, x, noutput=None):
    """digital_clock_recovery_mm_ff.general_work on the whole float stream x."""
    x = np.ascontiguousarray(x, np.float32)
    raw, ptr, view = aligned_stream(x, 1, np.float32)
    nout = noutput if noutput is not None else len(x)
    out = np.zeros(nout, np.float32)
    r = block.general_work(nout, [ptr], [len(x)], out)
    return out[:r], block.consumed


def run_pfb(block, x, numchans):
    """x: interleaved wideband stream (len multiple of numchans).  Performs
    gr_stream_to_streams, feeds the M history-prefixed streams, returns out[rows][M].
    The first general_work after set_taps returns 0 (gr_pfb_channelizer_ccf.cc:164-167)."""
    M = numchans
    x = np.ascontiguousarray(x, np.complex64)
    rows = len(x) // M
    xs = x[:rows * M].reshape(rows, M)
    hist = block.history
    keep, ptrs = [], []
    for j in range(M):
        raw, ptr, view = aligned_stream(xs[:, j], hist, np.complex64)
        keep.append(raw)
        ptrs.append(ptr)
    om = block.output_multiple
    rr = block.relative_rate  # 1/(M/os)
    nout = int(round(rows * rr * M))
    nout -= nout % om
    out = np.zeros(nout * M, np.complex64)
    r = block.general_work(nout, ptrs, [rows + hist - 1] * M, out)
    if r == 0:
        r = block.general_work(nout, ptrs, [rows + hist - 1] * M, out)
    return out[:r * M].reshape(r, M), block.consumed


# ---- scalar primitives -------------------------------------------------------------------
def fast_atan2f(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    o = np.zeros_like(y)
    lib().grref_fast_atan2f(_fp(y), _fp(x), _fp(o), C.c_long(len(y)))
    return o


def mmse_taps():
    o = np.zeros(129 * 8, np.float32)
    lib().grref_mmse_taps(_fp(o))
    return o.reshape(129, 8)


def rotator(incr, x):
    x = np.ascontiguousarray(x, np.complex64)
    o = np.zeros_like(x)
    lib().grref_rotator(C.c_float(np.float32(incr.real)), C.c_float(np.float32(incr.imag)),
                        x.ctypes.data_as(C.POINTER(C.c_float)), o.ctypes.data_as(C.POINTER(C.c_float)),
                        C.c_long(len(x)))
    return o


def _firdes(fn, *args):
    cap = 1 << 20
    o = np.zeros(cap, np.float32)
    n = fn(*args, _fp(o), cap)
    if n <= 0:
        raise RefError(lib().grref_last_error().decode())
    return o[:n].copy()


WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECTANGULAR, WIN_KAISER, WIN_BLACKMAN_HARRIS = range(6)


def firdes_low_pass(gain, fs, fc, tw, win=WIN_HAMMING, beta=6.76):
    return _firdes(lib().grref_firdes_low_pass, C.c_double(gain), C.c_double(fs), C.c_double(fc), C.c_double(tw),
                   int(win), C.c_double(beta))


def firdes_low_pass_2(gain, fs, fc, tw, atten, win=WIN_HAMMING, beta=6.76):
    return _firdes(lib().grref_firdes_low_pass_2, C.c_double(gain), C.c_double(fs), C.c_double(fc), C.c_double(tw),
                   C.c_double(atten), int(win), C.c_double(beta))


def firdes_root_raised_cosine(gain, fs, sym, alpha, ntaps):
    return _firdes(lib().grref_firdes_root_raised_cosine, C.c_double(gain), C.c_double(fs), C.c_double(sym),
                   C.c_double(alpha), int(ntaps))


def firdes_window(win, ntaps, beta=6.76):
    return _firdes(lib().grref_firdes_window, int(win), int(ntaps), C.c_double(beta))


# ---- flagship chain on the CPU, multi-threaded (bench.py --impl reference / cpu_baseline) --------
class _ChainParams(C.Structure):
    _fields_ = [("M", C.c_uint), ("pfb_taps", C.POINTER(C.c_float)), ("pfb_ntaps", C.c_int), ("quad_gain", C.c_float),
                ("rrc_taps", C.POINTER(C.c_float)), ("rrc_ntaps", C.c_int), ("omega", C.c_float),
                ("gain_omega", C.c_float), ("mu", C.c_float), ("gain_mu", C.c_float), ("limit", C.c_float),
                ("slicer_alpha", C.c_float), ("symbol_map", C.POINTER(C.c_int)), ("symbol_map_len", C.c_int),
                ("access_code", C.c_char_p), ("threshold", C.c_int)]


def bench_chain(M, pfb_taps, quad_gain, rrc_taps, omega, gain_omega, mu, gain_mu, limit, slicer_alpha, symbol_map,
                access_code, threshold, x, nthreads, fft_fast=True, keep_y=False):
    """Runs the reference's own blocks over x (rows*M interleaved complex64) with nthreads host threads.
    Returns (seconds_channelizer, seconds_demod, nhits, Y or None)."""
    L = lib()
    pt = np.ascontiguousarray(pfb_taps, np.float32)
    rt = np.ascontiguousarray(rrc_taps, np.float32)
    x = np.ascontiguousarray(x, np.complex64)
    rows = len(x) // M
    mp = (C.c_int * len(symbol_map))(*[int(v) for v in symbol_map])
    p = _ChainParams(M, _fp(pt), len(pt), quad_gain, _fp(rt), len(rt), omega, gain_omega, mu, gain_mu, limit,
                     slicer_alpha, C.cast(mp, C.POINTER(C.c_int)), len(symbol_map), access_code.encode(), threshold)
    secs = (C.c_double * 2)()
    nh = C.c_long(0)
    y = np.zeros(rows * M, np.complex64) if keep_y else None
    L.grref_set_fft_fast(int(bool(fft_fast)))
    try:
        rc = L.grref_bench_chain(C.byref(p), x.ctypes.data_as(C.POINTER(C.c_float)), C.c_long(rows), int(nthreads), secs,
                                 C.byref(nh), y.ctypes.data_as(C.POINTER(C.c_float)) if keep_y else None)
    finally:
        L.grref_set_fft_fast(0)
    assert rc == 0
    return secs[0], secs[1], nh.value, (y.reshape(rows, M) if keep_y else None)


def remez(order, bands, ampl, weight=(), filter_type="bandpass", grid_density=16):
    """The reference's gr_remez (general/gr_remez.cc:792-877), compiled in place.  Returns float64 taps or raises
    RuntimeError like gr.remez does."""
    b = np.ascontiguousarray(bands, np.float64)
    a = np.ascontiguousarray(ampl, np.float64)
    w = np.ascontiguousarray(weight if len(weight) else [], np.float64)
    out = np.zeros(order + 8, np.float64)
    L = lib()
    L.ref_remez.restype = C.c_int
    n = L.ref_remez(int(order), b.ctypes.data_as(C.POINTER(C.c_double)), len(b), a.ctypes.data_as(C.POINTER(C.c_double)),
                    w.ctypes.data_as(C.POINTER(C.c_double)), len(w), filter_type.encode(), int(grid_density),
                    out.ctypes.data_as(C.POINTER(C.c_double)))
    if n < 0:
        raise RuntimeError("gr_remez failed")
    return out[:n].copy()
