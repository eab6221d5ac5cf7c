"""Python mirror of the reference block interface for the hot path, over the C ABI of libgr_cuda.

Names, constructor arguments and error behaviour follow the reference's SWIG-exported blocks
(gr.fir_filter_ccf, gr.pfb_channelizer_ccf, digital.clock_recovery_mm_ff, ...; the SWIG magic
renames gr_make_X -> gr.X, gnuradio-core/src/lib/swig/gr_swig_block_magic.i:23-42):

  std::invalid_argument -> ValueError, std::out_of_range -> IndexError, CUDA failure -> GrCudaError.

`work()` takes numpy arrays laid out like the reference runtime hands them to a block: the input
starts history()-1 items before the first new item.  `work_device()` takes torch CUDA tensors
(or anything with .data_ptr()) and runs on the current torch stream.  `run()` plays
vector_source -> block -> vector_sink over a whole stream, keeping the history like gr_buffer.
"""
import ctypes as C

import numpy as np

from . import lib as _l
from .lib import ORDER_GENERIC, ORDER_SSE, Hit  # noqa: F401


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _dp(t):
    return C.c_void_p(t.data_ptr())


def _torch_stream():
    import torch
    s = torch.cuda.current_stream().cuda_stream
    return C.c_void_p(s if s else 1)   # torch's default stream (0) is cudaStreamLegacy (0x1); NULL = plan's own stream


class _Block:
    _destroy = None

    def __del__(self):
        try:
            if getattr(self, "h", None) and self._destroy:
                getattr(_l.load(), self._destroy)(self.h)
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------
class fir_filter_ccf(_Block):
    """gr_make_fir_filter_ccf(int decimation, const std::vector<float>& taps)
    (gnuradio-core/src/lib/filter/gr_fir_filter_XXX.cc.t:37-88)."""
    _destroy = "grcuda_fir_filter_ccf_destroy"
    in_dtype, out_dtype = np.complex64, np.complex64

    def __init__(self, decimation, taps):
        t = _f32(taps)
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_fir_filter_ccf_create(int(decimation), _p(t), len(t)))
        self._decim = int(decimation)

    def set_taps(self, taps):
        t = _f32(taps)
        _l.check(self.L.grcuda_fir_filter_ccf_set_taps(self.h, _p(t), len(t)))

    def history(self):
        return int(self.L.grcuda_fir_filter_ccf_history(self.h))

    def decimation(self):
        return self._decim

    def work(self, noutput_items, in_items):
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 0), np.complex64)
        n = _l.check(self.L.grcuda_fir_filter_ccf_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, noutput_items, d_in, d_out):
        _l.check(self.L.grcuda_fir_filter_ccf_work_device(self.h, C.c_long(noutput_items), _dp(d_in), _dp(d_out),
                                                          _torch_stream()))


class fir_filter_fff(_Block):
    """gr_make_fir_filter_fff(int decimation, const std::vector<float>& taps).  `order` selects which
    reference summation order is reproduced bit for bit (gr_fir_fff_simd.cc:99-134 vs generic)."""
    _destroy = "grcuda_fir_filter_fff_destroy"
    in_dtype, out_dtype = np.float32, np.float32

    def __init__(self, decimation, taps, order=ORDER_SSE):
        t = _f32(taps)
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_fir_filter_fff_create(int(decimation), _p(t), len(t), int(order)))
        self._decim = int(decimation)

    def set_taps(self, taps):
        t = _f32(taps)
        _l.check(self.L.grcuda_fir_filter_fff_set_taps(self.h, _p(t), len(t)))

    def history(self):
        return int(self.L.grcuda_fir_filter_fff_history(self.h))

    def decimation(self):
        return self._decim

    def work(self, noutput_items, in_items, abs_index0=None):
        x = _f32(in_items)
        out = np.empty(max(noutput_items, 0), np.float32)
        a0 = -(self.history() - 1) if abs_index0 is None else abs_index0
        n = _l.check(self.L.grcuda_fir_filter_fff_work(self.h, int(noutput_items), _p(x), _p(out), C.c_long(a0)))
        return out[:n]

    def work_device(self, noutput_items, nchan, d_in, d_out, abs_index0):
        _l.check(self.L.grcuda_fir_filter_fff_work_device(self.h, C.c_long(noutput_items), int(nchan), _dp(d_in),
                                                          _dp(d_out), C.c_long(abs_index0), _torch_stream()))


class freq_xlating_fir_filter_ccf(_Block):
    """gr_make_freq_xlating_fir_filter_ccf(int decimation, taps, double center_freq, double sampling_freq)
    (gr_freq_xlating_fir_filter_XXX.cc.t:38-123)."""
    _destroy = "grcuda_freq_xlating_fir_filter_ccf_destroy"
    in_dtype, out_dtype = np.complex64, np.complex64

    def __init__(self, decimation, taps, center_freq, sampling_freq):
        t = _f32(taps)
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_freq_xlating_fir_filter_ccf_create(
            int(decimation), _p(t), len(t), C.c_double(center_freq), C.c_double(sampling_freq)))
        self._decim = int(decimation)

    def set_taps(self, taps):
        t = _f32(taps)
        _l.check(self.L.grcuda_freq_xlating_fir_filter_ccf_set_taps(self.h, _p(t), len(t)))

    def set_center_freq(self, f):
        _l.check(self.L.grcuda_freq_xlating_fir_filter_ccf_set_center_freq(self.h, C.c_double(f)))

    def history(self):
        return int(self.L.grcuda_freq_xlating_fir_filter_ccf_history(self.h))

    def decimation(self):
        return self._decim

    def work(self, noutput_items, in_items):
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 0), np.complex64)
        n = _l.check(self.L.grcuda_freq_xlating_fir_filter_ccf_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, noutput_items, d_in, d_out):
        _l.check(self.L.grcuda_freq_xlating_fir_filter_ccf_work_device(self.h, C.c_long(noutput_items), _dp(d_in),
                                                                       _dp(d_out), _torch_stream()))


class pfb_channelizer_ccf(_Block):
    """gr_make_pfb_channelizer_ccf(unsigned numchans, taps, float oversample_rate=1)
    (gr_pfb_channelizer_ccf.cc:35-200).  ValueError if numchans/oversample_rate is not an integer."""
    _destroy = "grcuda_pfb_channelizer_ccf_destroy"

    def __init__(self, numchans, taps, oversample_rate=1.0):
        t = _f32(taps)
        self.L = _l.load()
        self.M = int(numchans)
        self.os = float(oversample_rate)
        self.h = _l.check_handle(self.L.grcuda_pfb_channelizer_ccf_create(self.M, _p(t), len(t), C.c_float(self.os)))

    def set_taps(self, taps):
        t = _f32(taps)
        _l.check(self.L.grcuda_pfb_channelizer_ccf_set_taps(self.h, _p(t), len(t)))

    def history(self):
        return int(self.L.grcuda_pfb_channelizer_ccf_history(self.h))

    def output_multiple(self):
        return int(self.L.grcuda_pfb_channelizer_ccf_output_multiple(self.h))

    def relative_rate(self):
        return float(self.L.grcuda_pfb_channelizer_ccf_relative_rate(self.h))

    def taps_per_filter(self):
        return int(self.L.grcuda_pfb_channelizer_ccf_taps_per_filter(self.h))

    def general_work(self, noutput_items, input_streams):
        """input_streams: list of numchans complex64 arrays (history-prefixed).  Returns (out[n][M], consumed)."""
        keep = [_c64(s) for s in input_streams]
        ptrs = (C.c_void_p * self.M)(*[s.ctypes.data for s in keep])
        out = np.empty((max(noutput_items, 0), self.M), np.complex64)
        consumed = C.c_int(0)
        n = _l.check(self.L.grcuda_pfb_channelizer_ccf_work(self.h, int(noutput_items), ptrs, _p(out), C.byref(consumed)))
        return out[:n], consumed.value

    def general_work_interleaved(self, noutput_items, rows):
        """rows: [(history-1) + nin][M] interleaved wideband samples (blks2.pfb_channelizer_ccf form)."""
        x = _c64(rows)
        out = np.empty((max(noutput_items, 0), self.M), np.complex64)
        consumed = C.c_int(0)
        n = _l.check(self.L.grcuda_pfb_channelizer_ccf_work_interleaved(self.h, int(noutput_items), _p(x), _p(out),
                                                                        C.byref(consumed)))
        return out[:n], consumed.value

    def work_device(self, noutput_items, d_in_rows, d_out):
        _l.check(self.L.grcuda_pfb_channelizer_ccf_work_device(self.h, C.c_long(noutput_items), _dp(d_in_rows),
                                                               _dp(d_out), _torch_stream()))


class pfb_arb_resampler_ccf(_Block):
    """gr_make_pfb_arb_resampler_ccf(float rate, taps, unsigned filter_size=32)
    (gr_pfb_arb_resampler_ccf.cc:42-205).  `nchan` > 1 is the batched [time][channel] form (one schedule for all
    channels); ValueError for fewer than 2 taps / a non-positive rate."""
    _destroy = "grcuda_pfb_arb_resampler_ccf_destroy"
    in_dtype, out_dtype = np.complex64, np.complex64

    def __init__(self, rate, taps, filter_size=32, nchan=1):
        t = _f32(taps)
        self.L = _l.load()
        self.nchan = int(nchan)
        self.h = _l.check_handle(self.L.grcuda_pfb_arb_resampler_ccf_create(C.c_float(rate), _p(t), len(t), int(filter_size),
                                                                            self.nchan))

    def set_rate(self, rate):
        _l.check(self.L.grcuda_pfb_arb_resampler_ccf_set_rate(self.h, C.c_float(rate)))

    def history(self):
        return int(self.L.grcuda_pfb_arb_resampler_ccf_history(self.h))

    def relative_rate(self):
        return float(self.L.grcuda_pfb_arb_resampler_ccf_relative_rate(self.h))

    def taps_per_filter(self):
        return int(self.L.grcuda_pfb_arb_resampler_ccf_taps_per_filter(self.h))

    def filter_taps(self, i, derivative=False):
        """print_taps() as data: the taps of polyphase filter i."""
        out = np.zeros(self.taps_per_filter(), np.float32)
        _l.check(self.L.grcuda_pfb_arb_resampler_ccf_get_taps(self.h, int(i), int(bool(derivative)), _p(out), len(out)))
        return out

    def general_work(self, noutput_items, in_items):
        """in_items: complex64 items from the first history item on.  Returns (out, consumed)."""
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 0), np.complex64)
        consumed = C.c_int(0)
        n = _l.check(self.L.grcuda_pfb_arb_resampler_ccf_work(self.h, int(noutput_items), len(x), _p(x), _p(out),
                                                              C.byref(consumed)))
        return out[:n], consumed.value

    def work_device(self, noutput_rows, ninput_rows, d_in, d_out):
        """d_in: [ninput_rows][nchan] complex64 on the device (row 0 = first history row); d_out: [>= noutput_rows][nchan].
        Returns (rows produced, rows consumed)."""
        consumed = C.c_int(0)
        n = _l.check(self.L.grcuda_pfb_arb_resampler_ccf_work_device(self.h, int(noutput_rows), int(ninput_rows), _dp(d_in),
                                                                     _dp(d_out), C.byref(consumed), _torch_stream()))
        return n, consumed.value

    def run(self, x, chunk_out=None):
        """vector_source -> block -> vector_sink over the new items x (zero history in front, like gr_buffer)."""
        x = _c64(x)
        buf = np.concatenate([np.zeros(self.history() - 1, np.complex64), x])
        rate = self.relative_rate()
        pos, outs, first = 0, [], True
        while True:
            nout = chunk_out or int((len(buf) - pos) * rate) + 16
            o, c = self.general_work(nout, buf[pos:])
            outs.append(o)
            pos += c
            if len(o) == 0 and c == 0:
                if first:
                    first = False
                    continue
                break
            first = False
        return np.concatenate(outs) if outs else np.zeros(0, np.complex64)


class pfb_decimator_ccf(_Block):
    """gr_make_pfb_decimator_ccf(unsigned decim, taps, unsigned channel) (gr_pfb_decimator_ccf.cc:35-175): the
    polyphase decimator that pulls ONE channel out of the wideband stream."""
    _destroy = "grcuda_pfb_decimator_ccf_destroy"

    def __init__(self, decim, taps, channel):
        t = _f32(taps)
        self.L = _l.load()
        self.M = int(decim)
        self.h = _l.check_handle(self.L.grcuda_pfb_decimator_ccf_create(self.M, _p(t), len(t), int(channel)))

    def set_taps(self, taps):
        t = _f32(taps)
        _l.check(self.L.grcuda_pfb_decimator_ccf_set_taps(self.h, _p(t), len(t)))

    def history(self):
        return int(self.L.grcuda_pfb_decimator_ccf_history(self.h))

    def taps_per_filter(self):
        return int(self.L.grcuda_pfb_decimator_ccf_taps_per_filter(self.h))

    def work(self, noutput_items, input_streams):
        """input_streams: decim complex64 arrays, each from its first history item.  Returns the outputs."""
        keep = [_c64(s) for s in input_streams]
        ptrs = (C.c_void_p * self.M)(*[s.ctypes.data for s in keep])
        out = np.empty(max(noutput_items, 0), np.complex64)
        n = _l.check(self.L.grcuda_pfb_decimator_ccf_work(self.h, int(noutput_items), ptrs, _p(out)))
        return out[:n]

    def work_interleaved(self, noutput_items, rows):
        x = _c64(rows)
        out = np.empty(max(noutput_items, 0), np.complex64)
        n = _l.check(self.L.grcuda_pfb_decimator_ccf_work_interleaved(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, noutput_items, d_in_rows, d_out):
        return _l.check(self.L.grcuda_pfb_decimator_ccf_work_device(self.h, C.c_long(noutput_items), _dp(d_in_rows), _dp(d_out),
                                                                    _torch_stream()))

    def run(self, x, chunk=None):
        """gr_stream_to_streams -> block -> vector_sink over the interleaved new items x (zero history in front)."""
        x = _c64(x)
        n = len(x) // self.M
        h = self.history() - 1
        rows = np.concatenate([np.zeros((h, self.M), np.complex64), x[:n * self.M].reshape(n, self.M)])
        assert self.work_interleaved(min(n, 4), rows).size == 0          # the "updated" call
        out, done, step = [], 0, chunk or max(n, 1)
        while done < n:
            m = min(step, n - done)
            out.append(self.work_interleaved(m, rows[done:done + m + h]))
            done += m
        return np.concatenate(out) if out else np.zeros(0, np.complex64)


class fft_filter_ccc(_Block):
    """gr_make_fft_filter_ccc(int decimation, const std::vector<gr_complex>& taps) (gr_fft_filter_ccc.cc:46-106)."""
    _destroy = "grcuda_fft_filter_ccc_destroy"
    in_dtype, out_dtype = np.complex64, np.complex64

    def __init__(self, decimation, taps):
        t = _c64(taps)
        self.L = _l.load()
        self.decim = int(decimation)
        self.h = _l.check_handle(self.L.grcuda_fft_filter_ccc_create(self.decim, _p(t), len(t)))

    def set_taps(self, taps):
        t = _c64(taps)
        _l.check(self.L.grcuda_fft_filter_ccc_set_taps(self.h, _p(t), len(t)))

    def output_multiple(self):
        return int(self.L.grcuda_fft_filter_ccc_output_multiple(self.h))

    def history(self):
        return 1

    def decimation(self):
        return self.decim

    def path(self):
        """0 direct form, 1 fused overlap-save (one CTA per block), 2 the same per 4096-tap partition (long filters)."""
        return int(self.L.grcuda_fft_filter_ccc_path(self.h))

    def set_path(self, path):
        """Pins a device path (-1: automatic); deferred like set_taps (the next work() returns 0)."""
        _l.check(self.L.grcuda_fft_filter_ccc_set_path(self.h, int(path)))

    def work(self, noutput_items, in_items):
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 0), np.complex64)
        n = _l.check(self.L.grcuda_fft_filter_ccc_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, noutput_items, d_in, d_out):
        return _l.check(self.L.grcuda_fft_filter_ccc_work_device(self.h, int(noutput_items), _dp(d_in), _dp(d_out),
                                                                 _torch_stream()))

    def run(self, x, blocks_per_call=None):
        """vector_source -> block -> vector_sink: whole blocks of output_multiple() items, like the scheduler."""
        x = _c64(x)
        ns = self.output_multiple()
        nblocks = (len(x) // self.decim) // ns
        outs, done, step = [], 0, (blocks_per_call or max(nblocks, 1)) * ns
        while done < nblocks * ns:
            n = min(step, nblocks * ns - done)
            outs.append(self.work(n, x[done * self.decim: (done + n) * self.decim]))
            done += n
        return np.concatenate(outs) if outs else np.zeros(0, np.complex64)


class fft_vcc(_Block):
    """gr_make_fft_vcc(int fft_size, bool forward, const std::vector<float>& window, bool shift=false)
    (gr_fft_vcc.cc:34-64, gr_fft_vcc_fftw.cc:51-103).  IndexError for fft_size <= 0."""
    _destroy = "grcuda_fft_vcc_destroy"

    def __init__(self, fft_size, forward, window, shift=False):
        w = _f32(window if window is not None else [])
        self.L = _l.load()
        self.n = int(fft_size)
        self.h = _l.check_handle(self.L.grcuda_fft_vcc_create(self.n, int(bool(forward)), _p(w), len(w), int(bool(shift))))

    def set_window(self, window):
        w = _f32(window)
        return bool(_l.check(self.L.grcuda_fft_vcc_set_window(self.h, _p(w), len(w))))

    def work(self, noutput_items, in_items):
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 0) * self.n, np.complex64)
        n = _l.check(self.L.grcuda_fft_vcc_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n * self.n]

    def work_device(self, noutput_items, d_in, d_out):
        _l.check(self.L.grcuda_fft_vcc_work_device(self.h, C.c_long(noutput_items), _dp(d_in), _dp(d_out), _torch_stream()))


class quadrature_demod_cf(_Block):
    """gr_make_quadrature_demod_cf(float gain) (gr_quadrature_demod_cf.cc:31-62); history 2."""
    _destroy = "grcuda_quadrature_demod_cf_destroy"
    in_dtype, out_dtype = np.complex64, np.float32

    def __init__(self, gain):
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_quadrature_demod_cf_create(C.c_float(gain)))

    def set_gain(self, gain):
        _l.check(self.L.grcuda_quadrature_demod_cf_set_gain(self.h, C.c_float(gain)))

    def gain(self):
        return float(self.L.grcuda_quadrature_demod_cf_gain(self.h))

    def history(self):
        return 2

    def decimation(self):
        return 1

    def work(self, noutput_items, in_items):
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 0), np.float32)
        n = _l.check(self.L.grcuda_quadrature_demod_cf_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, nrows, nchan, d_in, d_out):
        _l.check(self.L.grcuda_quadrature_demod_cf_work_device(self.h, C.c_long(nrows), int(nchan), _dp(d_in), _dp(d_out),
                                                               _torch_stream()))


def quad_demod_fir_fff_history(fir):
    """Rows of channelizer output the fused discriminator + matched filter wants in front of a block."""
    return int(_l.load().grcuda_quad_demod_fir_fff_history(fir.h))


def quad_demod_fir_fff_work_device(quad, fir, nrows, nchan, d_in, d_out, abs_row0):
    """quadrature_demod_cf -> fir_filter_fff fused into one kernel on [time][channel] data (SSE order):
    d_in = [history + nrows][nchan] complex64 CUDA tensor, d_out = [nrows][nchan] float32."""
    _l.check(_l.load().grcuda_quad_demod_fir_fff_work_device(quad.h, fir.h, C.c_long(nrows), int(nchan), _dp(d_in),
                                                            _dp(d_out), C.c_long(abs_row0), _torch_stream()))


class clock_recovery_mm_ff(_Block):
    """digital_make_clock_recovery_mm_ff(omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001)
    (gr-digital/lib/digital_clock_recovery_mm_ff.cc:36-139).  IndexError for omega < 1 or negative
    gains.  nchan > 1 gives the batched device form (independent loops per channel)."""
    _destroy = "grcuda_clock_recovery_mm_ff_destroy"

    def __init__(self, omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001, nchan=1, order=ORDER_SSE):
        self.L = _l.load()
        self.nchan = int(nchan)
        self.h = _l.check_handle(self.L.grcuda_clock_recovery_mm_ff_create(
            self.nchan, C.c_float(omega), C.c_float(gain_omega), C.c_float(mu), C.c_float(gain_mu),
            C.c_float(omega_relative_limit), int(order)))

    def forecast(self, noutput_items):
        return int(self.L.grcuda_clock_recovery_mm_ff_forecast(self.h, int(noutput_items)))

    def _state(self, chan=0):
        mu, om, last = C.c_float(), C.c_float(), C.c_float()
        _l.check(self.L.grcuda_clock_recovery_mm_ff_get_state(self.h, int(chan), C.byref(mu), C.byref(om), C.byref(last)))
        return mu.value, om.value, last.value

    def mu(self, chan=0):
        return self._state(chan)[0]

    def omega(self, chan=0):
        return self._state(chan)[1]

    def set_mu(self, mu):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_set_mu(self.h, C.c_float(mu)))

    def set_omega(self, omega):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_set_omega(self.h, C.c_float(omega)))

    def set_gain_mu(self, g):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_set_gain_mu(self.h, C.c_float(g)))

    def set_gain_omega(self, g):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_set_gain_omega(self.h, C.c_float(g)))

    def counters(self):
        a, b = C.c_longlong(), C.c_longlong()
        _l.check(self.L.grcuda_clock_recovery_mm_ff_counters(self.h, C.byref(a), C.byref(b)))
        return {"clamped": a.value, "overflow": b.value}

    def set_kernel_variant(self, variant):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_set_kernel_variant(self.h, int(variant)))

    def set_slicer(self, levels, alpha=0.0):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_set_slicer(self.h, int(levels), C.c_float(alpha)))

    def general_work(self, noutput_items, in_items, abs_index0=0):
        """Single-stream form: returns (out[:produced], consumed)."""
        x = _f32(in_items)
        out = np.empty(max(noutput_items, 1), np.float32)
        consumed = C.c_int(0)
        n = _l.check(self.L.grcuda_clock_recovery_mm_ff_work(self.h, int(noutput_items), len(x), _p(x), _p(out),
                                                             C.byref(consumed), C.c_long(abs_index0)))
        return out[:n], consumed.value

    def work_device(self, ninput_rows, abs_row0, d_in, d_out, d_slice, max_out, d_counts):
        _l.check(self.L.grcuda_clock_recovery_mm_ff_work_device(
            self.h, C.c_long(ninput_rows), C.c_long(abs_row0), _dp(d_in), _dp(d_out),
            _dp(d_slice) if d_slice is not None else None, int(max_out), _dp(d_counts), _torch_stream()))


class _slicer(_Block):
    _destroy = "grcuda_slicer_destroy"
    in_dtype, out_dtype = np.float32, np.uint8

    def history(self):
        return 1

    def decimation(self):
        return 1

    def work(self, noutput_items, in_items):
        x = _f32(in_items)
        out = np.empty(max(noutput_items, 0), np.uint8)
        n = _l.check(self.L.grcuda_slicer_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]


class pager_slicer_fb(_slicer):
    """pager_make_slicer_fb(float alpha) (gr-pager/lib/pager_slicer_fb.cc:29-84): the 4-level slicer."""

    def __init__(self, alpha):
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_pager_slicer_fb_create(C.c_float(alpha)))

    def dc_offset(self):
        return float(self.L.grcuda_pager_slicer_fb_dc_offset(self.h))


class binary_slicer_fb(_slicer):
    """digital_make_binary_slicer_fb() (gr-digital/lib/digital_binary_slicer_fb.cc:45-59)."""

    def __init__(self):
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_binary_slicer_fb_create())


class correlate_access_code_bb(_Block):
    """digital_make_correlate_access_code_bb(const std::string& access_code, int threshold)
    (gr-digital/lib/digital_correlate_access_code_bb.cc:36-133).  IndexError for > 64 bits."""
    _destroy = "grcuda_correlate_access_code_bb_destroy"
    in_dtype, out_dtype = np.uint8, np.uint8

    def __init__(self, access_code, threshold, nchan=1):
        self.L = _l.load()
        self.nchan = int(nchan)
        self.h = _l.check_handle(self.L.grcuda_correlate_access_code_bb_create(self.nchan, access_code.encode(), int(threshold)))

    def set_access_code(self, access_code):
        return _l.check(self.L.grcuda_correlate_access_code_bb_set_access_code(self.h, access_code.encode())) == 0

    def history(self):
        return 1

    def decimation(self):
        return 1

    def work(self, noutput_items, in_items):
        x = np.ascontiguousarray(in_items, np.uint8)
        out = np.empty(max(noutput_items, 0), np.uint8)
        n = _l.check(self.L.grcuda_correlate_access_code_bb_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_symbols_device(self, d_symbols, sym_rows, d_counts, symbol_map, bits_per_symbol, d_out, out_rows, d_hits,
                            max_hits, d_nhits):
        m = (C.c_int * len(symbol_map))(*[int(v) for v in symbol_map])
        _l.check(self.L.grcuda_correlate_access_code_bb_work_symbols_device(
            self.h, _dp(d_symbols), int(sym_rows), _dp(d_counts), m, len(symbol_map), int(bits_per_symbol),
            _dp(d_out) if d_out is not None else None, int(out_rows), _dp(d_hits), int(max_hits), _dp(d_nhits),
            _torch_stream()))


# ------------------------------------------------------------------------------------------------
def run(block, x, chunk=None, vlen=1):
    """vector_source -> sync block -> vector_sink over the new items x.  The harness keeps the
    history like gr_buffer (history-1 zero items pre-loaded, gr_buffer.cc:201-214) and honours the
    "work() returns 0 once after set_taps" contract."""
    x = np.ascontiguousarray(x, block.in_dtype)
    d = block.decimation()
    hist = block.history()
    buf = np.concatenate([np.zeros((hist - 1) * vlen, x.dtype), x])
    nout_total = (len(x) // vlen) // d
    outs, done = [], 0
    step = chunk or max(nout_total, 1)
    while done < nout_total:
        n = min(step, nout_total - done)
        y = block.work(n, buf[done * d * vlen:]) if not isinstance(block, fir_filter_fff) else \
            block.work(n, buf[done * d:], abs_index0=done * d - (hist - 1))
        if len(y) == 0 and n > 0:
            if block.history() != hist:
                hist = block.history()
                buf = np.concatenate([np.zeros((hist - 1) * vlen, x.dtype), x])
            continue
        outs.append(y.copy())
        done += len(y)
    return np.concatenate(outs) if outs else np.empty(0, block.out_dtype)


# ------------------------------------------------------------------------------------------------
def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


class map_bb(_Block):
    """gr_make_map_bb(const std::vector<int>& map) (gr_map_bb.cc:35-61)."""
    _destroy = "grcuda_map_bb_destroy"
    in_dtype, out_dtype = np.uint8, np.uint8

    def __init__(self, map):
        m = np.ascontiguousarray(map, dtype=np.int32)
        self.L = _l.load()
        self.h = _l.check_handle(self.L.grcuda_map_bb_create(_p(m), len(m)))

    def work(self, noutput_items, in_items):
        x = _u8(in_items)
        out = np.empty(max(noutput_items, 0), np.uint8)
        n = _l.check(self.L.grcuda_map_bb_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, noutput_items, d_in, d_out):
        return _l.check(self.L.grcuda_map_bb_work_device(self.h, C.c_long(int(noutput_items)), _dp(d_in), _dp(d_out), _torch_stream()))


class unpack_k_bits_bb(_Block):
    """gr_make_unpack_k_bits_bb(unsigned k) (gr_unpack_k_bits_bb.cc:38-70); IndexError (std::out_of_range) for k == 0."""
    _destroy = "grcuda_unpack_k_bits_bb_destroy"
    in_dtype, out_dtype = np.uint8, np.uint8

    def __init__(self, k):
        self.L = _l.load()
        self.k = int(k)
        self.h = _l.check_handle(self.L.grcuda_unpack_k_bits_bb_create(C.c_uint(self.k)))

    def interpolation(self):
        return int(self.L.grcuda_unpack_k_bits_bb_interpolation(self.h))

    def work(self, noutput_items, in_items):
        x = _u8(in_items)
        out = np.empty(max(noutput_items, 0), np.uint8)
        n = _l.check(self.L.grcuda_unpack_k_bits_bb_work(self.h, int(noutput_items), _p(x), _p(out)))
        return out[:n]

    def work_device(self, noutput_items, d_in, d_out):
        return _l.check(self.L.grcuda_unpack_k_bits_bb_work_device(self.h, C.c_long(int(noutput_items)), _dp(d_in), _dp(d_out),
                                                                   _torch_stream()))


class _streams(_Block):
    _destroy = "grcuda_streams_destroy"
    _create = None

    def __init__(self, item_size, nstreams):
        self.L = _l.load()
        self.item_size, self.nstreams = int(item_size), int(nstreams)
        self.h = _l.check_handle(getattr(self.L, self._create)(self.item_size, self.nstreams))

    def work(self, noutput_items, in_items):
        """in_items: noutput_items * nstreams items (any dtype of item_size bytes, or uint8 bytes).  Returns the list of
        nstreams output arrays (as bytes viewed back to the input dtype)."""
        x = np.ascontiguousarray(in_items)
        raw = x.view(np.uint8).reshape(-1)
        n = int(noutput_items)
        outs = [np.empty(n * self.item_size, np.uint8) for _ in range(self.nstreams)]
        ptrs = (C.c_void_p * self.nstreams)(*[o.ctypes.data for o in outs])
        _l.check(self.L.grcuda_streams_work(self.h, n, _p(raw), ptrs))
        if x.dtype.itemsize == self.item_size:
            return [o.view(x.dtype) for o in outs]
        return outs

    def work_device(self, noutput_items, d_in, d_out, out_stride_items=None):
        st = int(noutput_items) if out_stride_items is None else int(out_stride_items)
        return _l.check(self.L.grcuda_streams_work_device(self.h, C.c_long(int(noutput_items)), _dp(d_in), _dp(d_out), C.c_long(st),
                                                          _torch_stream()))


class stream_to_streams(_streams):
    """gr_make_stream_to_streams(size_t item_size, size_t nstreams) (gr_stream_to_streams.cc:37-66)."""
    _create = "grcuda_stream_to_streams_create"


class vector_to_streams(_streams):
    """gr_make_vector_to_streams(size_t item_size, size_t nstreams) (gr_vector_to_streams.cc:37-70)."""
    _create = "grcuda_vector_to_streams_create"


class framer_sink_1(_Block):
    """gr_make_framer_sink_1(gr_msg_queue_sptr target_queue) (gr_framer_sink_1.cc:75-196), batched over nchan streams.
    The target queue is the block's own device-side queue: messages() drains it as
    [(channel, whitener_offset, payload bytes)] in arrival order."""
    _destroy = "grcuda_framer_sink_1_destroy"

    def __init__(self, nchan=1, max_msgs=1 << 16, payload_capacity=1 << 24):
        self.L = _l.load()
        self.nchan, self.max_msgs, self.cap = int(nchan), int(max_msgs), int(payload_capacity)
        self.h = _l.check_handle(self.L.grcuda_framer_sink_1_create(self.nchan, self.max_msgs, self.cap))

    def work(self, in_items):
        x = _u8(in_items)
        return _l.check(self.L.grcuda_framer_sink_1_work(self.h, len(x), _p(x)))

    def work_device(self, nitems, d_in, item_stride, chan_stride, d_counts=None, count_scale=1):
        return _l.check(self.L.grcuda_framer_sink_1_work_device(
            self.h, C.c_long(int(nitems)), _dp(d_in), C.c_long(int(item_stride)), C.c_long(int(chan_stride)),
            _dp(d_counts) if d_counts is not None else None, int(count_scale), _torch_stream()))

    def messages(self, with_meta=False):
        msgs = (_l.FramerMsg * self.max_msgs)()
        payload = np.empty(self.cap, np.uint8)
        dropped = C.c_int(0)
        n = _l.check(self.L.grcuda_framer_sink_1_read(self.h, msgs, self.max_msgs, _p(payload), C.c_size_t(self.cap), C.byref(dropped)))
        self.dropped = dropped.value
        out = []
        for m in msgs[:n]:
            data = bytes(payload[m.payload_offset: m.payload_offset + m.length]) if m.payload_offset >= 0 else None
            out.append((m.channel, m.whitener_offset, data, m.end_index, m.seq) if with_meta else (m.channel, m.whitener_offset, data))
        return out


class clock_recovery_mm_cc(_Block):
    """digital_make_clock_recovery_mm_cc(omega, gain_omega, mu, gain_mu, omega_relative_limit)
    (digital_clock_recovery_mm_cc.cc:37-75), + nchan for the batched device form.  IndexError (std::out_of_range) for
    omega <= 0 or negative gains."""
    _destroy = "grcuda_clock_recovery_mm_cc_destroy"

    def __init__(self, omega, gain_omega, mu, gain_mu, omega_relative_limit=0.001, nchan=1):
        self.L = _l.load()
        self.nchan = int(nchan)
        self.h = _l.check_handle(self.L.grcuda_clock_recovery_mm_cc_create(
            self.nchan, C.c_float(omega), C.c_float(gain_omega), C.c_float(mu), C.c_float(gain_mu), C.c_float(omega_relative_limit)))

    def history(self):
        return 3

    def forecast(self, noutput_items):
        return _l.check(self.L.grcuda_clock_recovery_mm_cc_forecast(self.h, int(noutput_items)))

    def state(self, chan=0):
        mu, om = C.c_float(0), C.c_float(0)
        _l.check(self.L.grcuda_clock_recovery_mm_cc_get_state(self.h, int(chan), C.byref(mu), C.byref(om)))
        return mu.value, om.value

    def mu(self):
        return self.state()[0]

    def omega(self):
        return self.state()[1]

    def set_mu(self, mu):
        _l.check(self.L.grcuda_clock_recovery_mm_cc_set_mu(self.h, C.c_float(mu)))

    def set_omega(self, omega):
        _l.check(self.L.grcuda_clock_recovery_mm_cc_set_omega(self.h, C.c_float(omega)))

    def set_gain_mu(self, g):
        _l.check(self.L.grcuda_clock_recovery_mm_cc_set_gain_mu(self.h, C.c_float(g)))

    def set_gain_omega(self, g):
        _l.check(self.L.grcuda_clock_recovery_mm_cc_set_gain_omega(self.h, C.c_float(g)))

    def counters(self):
        a, b = C.c_longlong(0), C.c_longlong(0)
        _l.check(self.L.grcuda_clock_recovery_mm_cc_counters(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def general_work(self, noutput_items, in_items, with_error=False):
        """Returns (symbols, error signal or None, consumed)."""
        x = _c64(in_items)
        out = np.empty(max(noutput_items, 1), np.complex64)
        err = np.empty(max(noutput_items, 1), np.float32) if with_error else None
        consumed = C.c_int(0)
        n = _l.check(self.L.grcuda_clock_recovery_mm_cc_work(self.h, int(noutput_items), len(x), _p(x), _p(out),
                                                             _p(err) if with_error else None, C.byref(consumed)))
        return out[:n], (err[:n] if with_error else None), consumed.value

    def work_device(self, ninput_rows, abs_row0, d_in, d_out, d_err, max_out, d_counts):
        return _l.check(self.L.grcuda_clock_recovery_mm_cc_work_device(
            self.h, C.c_long(int(ninput_rows)), C.c_long(int(abs_row0)), _dp(d_in), _dp(d_out),
            _dp(d_err) if d_err is not None else None, int(max_out), _dp(d_counts) if d_counts is not None else None, _torch_stream()))
