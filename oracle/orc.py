"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product package (grb200).

numpy/ctypes front end of oracle/liboracle.so (oracle.c: the plain-C restatement of the
reference algorithms).  Stream helpers take NEW items only and prepend the zero history the
reference runtime pre-loads (gr_buffer.cc:201-214)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
ORDER_GENERIC, ORDER_SSE = 0, 1


class MMState(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("mu", "omega", "min_omega", "omega_mid", "max_omega", "gain_omega",
                                         "gain_mu", "last_sample", "omega_relative_limit")]


class Slicer4State(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("avg", C.c_float)]


class CorrState(C.Structure):
    _fields_ = [(n, C.c_ulonglong) for n in ("access_code", "data_reg", "flag_reg", "flag_bit", "mask")] + \
               [("threshold", C.c_uint)]


class ArbState(C.Structure):
    _fields_ = [("int_rate", C.c_uint), ("dec_rate", C.c_uint), ("last_filter", C.c_uint), ("taps_per_filter", C.c_uint),
                ("flt_rate", C.c_float), ("acc", C.c_float), ("rate", C.c_float), ("start_index", C.c_int),
                ("updated", C.c_int), ("taps", C.POINTER(C.c_float)), ("dtaps", C.POINTER(C.c_float))]


class FftFiltState(C.Structure):
    _fields_ = [("ntaps", C.c_int), ("fftsize", C.c_int), ("nsamples", C.c_int), ("decimation", C.c_int),
                ("xformed_taps", C.c_void_p), ("tail", C.c_void_p)]


class Rotator(C.Structure):
    _fields_ = [("phase", C.c_float * 2), ("incr", C.c_float * 2), ("counter", C.c_uint)]


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_fast_atan2f.restype = C.c_float
        L.orc_fast_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_mmse_interpolate.restype = C.c_float
        L.orc_mmse_table.restype = C.POINTER(C.c_float)
        L.orc_atan_table.restype = C.POINTER(C.c_float)
        L.orc_count_bits64.argtypes = [C.c_ulonglong]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _hist(x, h, dtype):
    x = np.ascontiguousarray(x, dtype)
    return np.concatenate([np.zeros(h, dtype), x])


def fir_ccf(taps, decim, x, hist_prefixed=False):
    t = np.ascontiguousarray(taps, np.float32)
    xin = np.ascontiguousarray(x, np.complex64) if hist_prefixed else _hist(x, max(len(t) - 1, 0), np.complex64)
    nout = (len(xin) - max(len(t) - 1, 0)) // decim
    out = np.zeros(nout, np.complex64)
    lib().orc_fir_ccf(_p(t), len(t), int(decim), _p(xin), C.c_long(nout), _p(out))
    return out


def fir_fff(taps, decim, x, order=ORDER_GENERIC, abs0=None, hist_prefixed=False):
    t = np.ascontiguousarray(taps, np.float32)
    h = max(len(t) - 1, 0)
    xin = np.ascontiguousarray(x, np.float32) if hist_prefixed else _hist(x, h, np.float32)
    nout = (len(xin) - h) // decim
    out = np.zeros(nout, np.float32)
    a0 = -h if abs0 is None else abs0
    lib().orc_fir_fff(_p(t), len(t), int(decim), _p(xin), C.c_long(nout), _p(out), int(order), C.c_long(a0))
    return out


def fir_ccc(taps, decim, x):
    t = np.ascontiguousarray(taps, np.complex64)
    xin = _hist(x, max(len(t) - 1, 0), np.complex64)
    nout = (len(xin) - max(len(t) - 1, 0)) // decim
    out = np.zeros(nout, np.complex64)
    lib().orc_fir_ccc(_p(t), len(t), int(decim), _p(xin), C.c_long(nout), _p(out))
    return out


def rotator_new(incr):
    r = Rotator()
    lib().orc_rotator_init_f(C.byref(r), C.c_float(np.float32(incr.real)), C.c_float(np.float32(incr.imag)))
    return r


def rotate(incr, x):
    r = rotator_new(incr)
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    lib().orc_rotate_n(C.byref(r), _p(x), C.c_long(len(x)), _p(out))
    return out


def freq_xlating_taps(proto, center_freq, sampling_freq, decim):
    p = np.ascontiguousarray(proto, np.float32)
    ct = np.zeros(len(p), np.complex64)
    inc = np.zeros(1, np.complex64)
    lib().orc_freq_xlating_taps(_p(p), len(p), C.c_double(center_freq), C.c_double(sampling_freq), int(decim),
                                _p(ct), _p(inc))
    return ct, inc[0]


def freq_xlating_fir_ccf(proto, decim, center_freq, sampling_freq, x, rot=None):
    p = np.ascontiguousarray(proto, np.float32)
    xin = _hist(x, max(len(p) - 1, 0), np.complex64)
    nout = (len(xin) - max(len(p) - 1, 0)) // decim
    out = np.zeros(nout, np.complex64)
    if rot is None:
        _, inc = freq_xlating_taps(p, center_freq, sampling_freq, decim)
        rot = rotator_new(complex(inc))
    lib().orc_freq_xlating_fir_ccf(_p(p), len(p), int(decim), C.c_double(center_freq), C.c_double(sampling_freq),
                                   C.byref(rot), _p(xin), C.c_long(nout), _p(out))
    return out


def dft(x, forward=True):
    x = np.ascontiguousarray(x, np.complex64)
    out = np.zeros_like(x)
    lib().orc_dft(_p(x), _p(out), len(x), int(bool(forward)))
    return out


def pfb_taps_per_filter(M, ntaps):
    return lib().orc_pfb_taps_per_filter(int(M), int(ntaps))


def pfb_output_multiple(M, os_rate):
    return lib().orc_pfb_output_multiple(int(M), C.c_float(os_rate))


def pfb_check_rate(M, os_rate):
    return bool(lib().orc_pfb_check_rate(int(M), C.c_float(os_rate)))


def pfb_channelizer_ccf(M, taps, x, oversample_rate=1.0):
    """x: interleaved wideband stream; returns (out[rows_out][M], consumed_per_stream)."""
    t = np.ascontiguousarray(taps, np.float32)
    T = pfb_taps_per_filter(M, len(t))
    x = np.ascontiguousarray(x, np.complex64)
    rows = len(x) // M
    xs = x[:rows * M].reshape(rows, M)
    streams = [np.ascontiguousarray(np.concatenate([np.zeros(T, np.complex64), xs[:, j]])) for j in range(M)]
    ptrs = (C.c_void_p * M)(*[s.ctypes.data for s in streams])
    om = pfb_output_multiple(M, oversample_rate)
    nout = int(round(rows * oversample_rate))
    nout -= nout % om
    out = np.zeros(nout * M, np.complex64)
    consumed = C.c_int(0)
    r = lib().orc_pfb_channelizer_ccf(int(M), _p(t), len(t), C.c_float(oversample_rate), ptrs, nout, _p(out),
                                      C.byref(consumed))
    return out.reshape(r, M), consumed.value


def fft_vcc(N, forward, window, shift, x):
    x = np.ascontiguousarray(x, np.complex64)
    w = np.ascontiguousarray(window if window is not None else [], np.float32)
    nvec = len(x) // N
    out = np.zeros(nvec * N, np.complex64)
    lib().orc_fft_vcc(int(N), int(bool(forward)), _p(w) if len(w) else None, len(w), int(bool(shift)), _p(x),
                      C.c_long(nvec), _p(out))
    return out


def fast_atan2f(y, x):
    y = np.ascontiguousarray(y, np.float32)
    x = np.ascontiguousarray(x, np.float32)
    f = lib().orc_fast_atan2f
    return np.array([f(float(a), float(b)) for a, b in zip(y, x)], np.float32)


def quadrature_demod_cf(gain, x, prev=0j):
    xin = np.concatenate([np.array([prev], np.complex64), np.ascontiguousarray(x, np.complex64)])
    out = np.zeros(len(xin) - 1, np.float32)
    lib().orc_quadrature_demod_cf(C.c_float(gain), _p(xin), C.c_long(len(out)), _p(out))
    return out


def mmse_interpolate(in8, mu, order=ORDER_GENERIC, abs0=0):
    a = np.ascontiguousarray(in8, np.float32)
    return lib().orc_mmse_interpolate(_p(a), C.c_float(mu), int(order), C.c_long(abs0))


def mm_new(omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001):
    s = MMState()
    if lib().orc_mm_init(C.byref(s), C.c_float(omega), C.c_float(gain_omega), C.c_float(mu), C.c_float(gain_mu),
                         C.c_float(omega_relative_limit)) != 0:
        raise IndexError("out_of_range")
    return s


def mm_work(state, x, noutput=None, order=ORDER_GENERIC, abs0=0):
    x = np.ascontiguousarray(x, np.float32)
    nout = len(x) if noutput is None else noutput
    out = np.zeros(max(nout, 1), np.float32)
    consumed = C.c_int(0)
    r = lib().orc_mm_general_work(C.byref(state), _p(x), len(x), _p(out), int(nout), C.byref(consumed), int(order),
                                  C.c_long(abs0))
    return out[:r], consumed.value


def slicer4(x, alpha=0.0, state=None):
    s = state or Slicer4State()
    if state is None:
        lib().orc_slicer4_init(C.byref(s), C.c_float(alpha))
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x), np.uint8)
    lib().orc_slicer4(C.byref(s), _p(x), C.c_long(len(x)), _p(out))
    return out


def binary_slicer(x):
    x = np.ascontiguousarray(x, np.float32)
    out = np.zeros(len(x), np.uint8)
    lib().orc_binary_slicer(_p(x), C.c_long(len(x)), _p(out))
    return out


def map_bb(m, x):
    x = np.ascontiguousarray(x, np.uint8)
    mm = (C.c_int * len(m))(*[int(v) for v in m])
    out = np.zeros(len(x), np.uint8)
    lib().orc_map_bb(mm, len(m), _p(x), C.c_long(len(x)), _p(out))
    return out


def unpack_k_bits_bb(k, x):
    x = np.ascontiguousarray(x, np.uint8)
    out = np.zeros(len(x) * k, np.uint8)
    lib().orc_unpack_k_bits_bb(int(k), _p(x), C.c_long(len(x)), _p(out))
    return out


def corr_new(code, threshold):
    s = CorrState()
    if lib().orc_corr_init(C.byref(s), code.encode(), int(threshold)) != 0:
        raise IndexError("out_of_range: access_code is > 64 bits")
    return s


def corr_work(state, bits):
    b = np.ascontiguousarray(bits, np.uint8)
    out = np.zeros(len(b), np.uint8)
    lib().orc_corr_work(C.byref(state), _p(b), C.c_long(len(b)), _p(out))
    return out


def firdes_window(win, ntaps, beta=6.76):
    out = np.zeros(ntaps, np.float32)
    if lib().orc_firdes_window(int(win), int(ntaps), C.c_double(beta), _p(out)) < 0:
        raise IndexError("gr_firdes:window: type out of range")
    return out


def firdes_low_pass(gain, fs, fc, tw, win=0, beta=6.76):
    out = np.zeros(1 << 20, np.float32)
    n = lib().orc_firdes_low_pass(C.c_double(gain), C.c_double(fs), C.c_double(fc), C.c_double(tw), int(win),
                                  C.c_double(beta), _p(out), len(out))
    if n <= 0:
        raise ValueError("firdes check failed")
    return out[:n].copy()


def firdes_low_pass_2(gain, fs, fc, tw, atten, win=0, beta=6.76):
    out = np.zeros(1 << 20, np.float32)
    n = lib().orc_firdes_low_pass_2(C.c_double(gain), C.c_double(fs), C.c_double(fc), C.c_double(tw),
                                    C.c_double(atten), int(win), C.c_double(beta), _p(out), len(out))
    if n <= 0:
        raise ValueError("firdes check failed")
    return out[:n].copy()


def firdes_root_raised_cosine(gain, fs, sym, alpha, ntaps):
    out = np.zeros(ntaps | 1, np.float32)
    n = lib().orc_firdes_root_raised_cosine(C.c_double(gain), C.c_double(fs), C.c_double(sym), C.c_double(alpha),
                                            int(ntaps), _p(out))
    return out[:n]


def mmse_table():
    return np.ctypeslib.as_array(lib().orc_mmse_table(), shape=(129, 8)).copy()


def atan_table():
    return np.ctypeslib.as_array(lib().orc_atan_table(), shape=(257,)).copy()


# ---- gr_pfb_arb_resampler_ccf -----------------------------------------------------------------------
class ArbResampler:
    """orc_arb_* (gr_pfb_arb_resampler_ccf.cc): a stateful block; run() plays the scheduler around it."""

    def __init__(self, rate, taps, filter_size=32):
        t = np.ascontiguousarray(taps, np.float32)
        self.s = ArbState()
        if lib().orc_arb_init(C.byref(self.s), C.c_float(rate), _p(t), len(t), int(filter_size)) != 0:
            raise ValueError("pfb_arb_resampler_ccf: needs at least 2 taps and 1 filter")

    def __del__(self):
        try:
            lib().orc_arb_free(C.byref(self.s))
        except Exception:
            pass

    history = property(lambda self: int(self.s.taps_per_filter) + 1)

    def set_rate(self, rate):
        lib().orc_arb_set_rate(C.byref(self.s), C.c_float(rate))

    def filter_taps(self, i):
        T = int(self.s.taps_per_filter)
        return np.array([self.s.taps[i * T + j] for j in range(T)], np.float32)

    def general_work(self, noutput, in_items):
        """in_items starts at the first history item.  Returns (out, consumed)."""
        x = np.ascontiguousarray(in_items, np.complex64)
        out = np.zeros(max(noutput, 1), np.complex64)
        consumed = C.c_int(0)
        r = lib().orc_arb_general_work(C.byref(self.s), _p(x), len(x), _p(out), int(noutput), C.byref(consumed))
        return out[:r], consumed.value

    def schedule(self, ninput, noutput):
        cnt = np.zeros(max(noutput, 1), np.int32)
        flt = np.zeros(max(noutput, 1), np.uint16)
        acc = np.zeros(max(noutput, 1), np.float32)
        consumed = C.c_int(0)
        r = lib().orc_arb_schedule(C.byref(self.s), int(ninput), int(noutput), _p(cnt), _p(flt), _p(acc), C.byref(consumed))
        return cnt[:r], flt[:r], acc[:r], consumed.value

    def run(self, x, chunk_out=None):
        """vector_source -> block -> vector_sink over the new items x (zero history in front)."""
        return run_general(self, x, chunk_out, float(self.s.rate))


def run_general(block, x, chunk_out, rate):
    """Scheduler loop shared with oracle/refharness.py's run_arb: `block.general_work(noutput, items from the first
    history item)` -> (out, consumed); stops when a call neither produces nor consumes (after the first, which
    returns 0 by contract)."""
    x = np.ascontiguousarray(x, np.complex64)
    h = block.history - 1
    buf = np.concatenate([np.zeros(h, np.complex64), x])
    pos, outs, first = 0, [], True
    while True:
        avail = len(buf) - pos
        nout = chunk_out or int(avail * rate) + 16
        o, c = block.general_work(nout, buf[pos:])
        outs.append(o)
        pos += c
        if len(o) == 0 and c == 0:
            if first:
                first = False
                continue
            break
        first = False
    return np.concatenate(outs) if outs else np.zeros(0, np.complex64)


# ---- gr_pfb_decimator_ccf ---------------------------------------------------------------------------
def pfb_decimator_ccf(decim, taps, channel, x):
    """x: interleaved stream (len multiple of decim; stream s = x[m*decim + s], gr_stream_to_streams).  Zero history
    (taps_per_filter - 1 items per stream) is prepended.  Returns the len(x)/decim outputs."""
    t = np.ascontiguousarray(taps, np.float32)
    x = np.ascontiguousarray(x, np.complex64)
    T = lib().orc_pfb_decimator_taps_per_filter(int(decim), len(t))
    n = len(x) // decim
    streams = [np.concatenate([np.zeros(T - 1, np.complex64), x[s::decim][:n]]) for s in range(decim)]
    ptrs = (C.c_void_p * decim)(*[s.ctypes.data for s in streams])
    out = np.zeros(max(n, 1), np.complex64)
    lib().orc_pfb_decimator_ccf(int(decim), _p(t), len(t), int(channel), ptrs, C.c_long(n), _p(out))
    return out[:n]


# ---- gr_fft_filter_ccc ------------------------------------------------------------------------------
class FftFilter:
    """orc_fftfilt_*: overlap-add FFT filter with complex taps (gri_fft_filter_ccc_generic.cc)."""

    def __init__(self, decim, taps):
        t = np.ascontiguousarray(taps, np.complex64)
        self.s = FftFiltState()
        self.decim = int(decim)
        self.nsamples = lib().orc_fftfilt_init(C.byref(self.s), self.decim, _p(t), len(t))

    def __del__(self):
        try:
            lib().orc_fftfilt_free(C.byref(self.s))
        except Exception:
            pass

    def set_taps(self, taps):
        t = np.ascontiguousarray(taps, np.complex64)
        self.nsamples = lib().orc_fftfilt_set_taps(C.byref(self.s), _p(t), len(t))

    def filter(self, nitems, x):
        x = np.ascontiguousarray(x, np.complex64)
        assert nitems % self.nsamples == 0 and len(x) >= nitems * self.decim
        out = np.zeros(max(nitems, 1), np.complex64)
        lib().orc_fftfilt_filter(C.byref(self.s), int(nitems), _p(x), _p(out))
        return out[:nitems]

    def run(self, x, blocks_per_call=None):
        x = np.ascontiguousarray(x, np.complex64)
        ns = self.nsamples
        nblocks = (len(x) // self.decim) // ns
        outs, done, step = [], 0, (blocks_per_call or max(nblocks, 1)) * ns
        while done < nblocks * ns:
            n = min(step, nblocks * ns - done)
            outs.append(self.filter(n, x[done * self.decim:]))
            done += n
        return np.concatenate(outs) if outs else np.zeros(0, np.complex64)


# ---- gr_framer_sink_1 ---------------------------------------------------------------------------------
class FramerState(C.Structure):
    _fields_ = [("state", C.c_int), ("header", C.c_uint), ("headerbitlen_cnt", C.c_int), ("packetlen", C.c_int),
                ("whitener_offset", C.c_int), ("packetlen_cnt", C.c_int), ("byte_index", C.c_int),
                ("packet_byte", C.c_ubyte), ("packet", C.c_ubyte * 4096)]


_EMIT = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.POINTER(C.c_ubyte), C.c_int)


class Framer:
    """orc_framer_*: gr_framer_sink_1's state machine; work() returns the packets posted during the call as
    [(whitener_offset, payload bytes)]."""

    def __init__(self):
        self.s = FramerState()
        lib().orc_framer_init(C.byref(self.s))

    def work(self, in_bytes):
        x = np.ascontiguousarray(in_bytes, np.uint8)
        got = []

        def emit(ctx, off, payload, n):
            got.append((int(off), bytes(bytearray(payload[i] for i in range(n)))))
        cb = _EMIT(emit)
        lib().orc_framer_work(C.byref(self.s), _p(x), C.c_long(len(x)), cb, None)
        return got


def framer_make_stream(rng, packets, gap=(5, 200), corrupt_header_every=0):
    """Correlator-style byte stream (bit 0 data, bit 1 flag) carrying `packets` = [(whitener_offset, payload bytes)]:
    random data bits, then for each packet one byte with the flag set (the first header bit rides on it), 32 header bits
    and the payload MSB first.  corrupt_header_every = k flips one header bit of every k-th packet."""
    out = []
    for i, (off, payload) in enumerate(packets):
        out.append(rng.integers(0, 2, int(rng.integers(gap[0], gap[1]))).astype(np.uint8))
        h16 = ((off & 0xf) << 12) | (len(payload) & 0xfff)
        hdr = [(((h16 << 16) | h16) >> (31 - b)) & 1 for b in range(32)]
        if corrupt_header_every and (i + 1) % corrupt_header_every == 0:
            hdr[int(rng.integers(0, 32))] ^= 1
        bits = np.array(hdr + [(byte >> (7 - b)) & 1 for byte in payload for b in range(8)], np.uint8)
        bits[0] |= 2
        out.append(bits)
    out.append(rng.integers(0, 2, 50).astype(np.uint8))
    return np.concatenate(out)


# ---- digital_clock_recovery_mm_cc ---------------------------------------------------------------------
class MMCCState(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("mu", "omega", "gain_omega", "gain_mu", "omega_mid", "omega_relative_limit")] + \
               [(n, C.c_float * 2) for n in ("p_2T", "p_1T", "p_0T", "c_2T", "c_1T", "c_0T")]


def mmcc_new(omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001):
    s = MMCCState()
    if lib().orc_mmcc_init(C.byref(s), C.c_float(omega), C.c_float(gain_omega), C.c_float(mu), C.c_float(gain_mu),
                           C.c_float(omega_relative_limit)) != 0:
        raise IndexError("out_of_range")
    return s


def mmcc_work(state, x, noutput=None, with_error=False):
    """Returns (symbols, error signal or None, consumed)."""
    x = np.ascontiguousarray(x, np.complex64)
    nout = len(x) if noutput is None else noutput
    out = np.zeros(max(nout, 1), np.complex64)
    err = np.zeros(max(nout, 1), np.float32) if with_error else None
    consumed = C.c_int(0)
    r = lib().orc_mmcc_general_work(C.byref(state), _p(x), len(x), _p(out), _p(err) if with_error else None, int(nout),
                                    C.byref(consumed))
    return out[:r], (err[:r] if with_error else None), consumed.value
