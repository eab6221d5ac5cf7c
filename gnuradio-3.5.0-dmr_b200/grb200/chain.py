"""The flagship pipeline object (include/gr_cuda.h: grcuda_dmr_chain_*): wideband interleaved stream ->
PFB channelizer -> batched 4FSK demod -> sync search, all M channels at once, HBM resident."""
import ctypes as C

import numpy as np

from . import lib as _l
from . import synth


def _sp(stream):
    """None -> the plan's own stream; 0 (torch's default stream) -> cudaStreamLegacy; else the handle."""
    if stream is None:
        return None
    return C.c_void_p(stream if stream else 1)


class DmrChainConfig:
    """Parameters of the reference flowgraph this object stands for."""

    def __init__(self, numchans, pfb_taps, fs_channel=12500.0, rrc_taps=None, rrc_ntaps=29, omega=None, gain_mu=0.175,
                 gain_omega=None, mu=0.5, omega_relative_limit=0.005, slicer_alpha=0.0, symbol_map=None,
                 access_code=None, threshold=2, order=_l.ORDER_SSE, max_rows_per_block=4096, keep_bytes=False,
                 quad_gain=None):
        self.numchans = int(numchans)
        self.pfb_taps = np.ascontiguousarray(pfb_taps, np.float32)
        self.fs_channel = float(fs_channel)
        self.quad_gain = float(quad_gain if quad_gain is not None else fs_channel / (2 * np.pi * synth.DEVIATION_HZ))
        if rrc_taps is None:
            from . import firdes
            rrc_taps = firdes.root_raised_cosine(1.0, fs_channel, synth.SYMBOL_RATE, synth.RRC_ALPHA, rrc_ntaps)
        self.rrc_taps = np.ascontiguousarray(rrc_taps, np.float32)
        self.omega = float(omega if omega is not None else fs_channel / synth.SYMBOL_RATE)
        self.gain_mu = float(gain_mu)
        self.gain_omega = float(gain_omega if gain_omega is not None else 0.25 * gain_mu * gain_mu)
        self.mu = float(mu)
        self.omega_relative_limit = float(omega_relative_limit)
        self.slicer_alpha = float(slicer_alpha)
        self.symbol_map = list(symbol_map if symbol_map is not None else synth.SLICER_TO_DIBIT_MAP)
        self.access_code = access_code if access_code is not None else synth.access_code_string(synth.DMR_BS_DATA_SYNC_BITS)
        self.threshold = int(threshold)
        self.order = int(order)
        self.max_rows_per_block = int(max_rows_per_block)
        self.keep_bytes = bool(keep_bytes)


class DmrChain:
    def __init__(self, cfg):
        self.cfg = cfg
        self.L = _l.load()
        p = _l.ChainParams()
        p.numchans = cfg.numchans
        p.pfb_taps = cfg.pfb_taps.ctypes.data_as(C.POINTER(C.c_float))
        p.pfb_ntaps = len(cfg.pfb_taps)
        p.quad_gain = cfg.quad_gain
        p.rrc_taps = cfg.rrc_taps.ctypes.data_as(C.POINTER(C.c_float))
        p.rrc_ntaps = len(cfg.rrc_taps)
        p.omega, p.gain_omega, p.mu, p.gain_mu = cfg.omega, cfg.gain_omega, cfg.mu, cfg.gain_mu
        p.omega_relative_limit = cfg.omega_relative_limit
        p.slicer_alpha = cfg.slicer_alpha
        self._map = (C.c_int * len(cfg.symbol_map))(*cfg.symbol_map)
        p.symbol_map = C.cast(self._map, C.POINTER(C.c_int))
        p.symbol_map_len = len(cfg.symbol_map)
        p.access_code = cfg.access_code.encode()
        p.threshold = cfg.threshold
        p.order = cfg.order
        p.max_rows_per_block = cfg.max_rows_per_block
        p.keep_bytes = int(cfg.keep_bytes)
        self.h = _l.check_handle(self.L.grcuda_dmr_chain_create(C.byref(p)))
        self.M = cfg.numchans

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.grcuda_dmr_chain_destroy(self.h)
        except Exception:
            pass

    def history_rows(self):
        return int(self.L.grcuda_dmr_chain_history_rows(self.h))

    def min_rows(self):
        return int(self.L.grcuda_dmr_chain_min_rows(self.h))

    def warmup_rows(self):
        return int(self.L.grcuda_dmr_chain_warmup_rows(self.h))

    def seek(self, abs_row):
        _l.check(self.L.grcuda_dmr_chain_seek(self.h, C.c_longlong(abs_row)))

    def seek_async(self, abs_row, stream=None):
        _l.check(self.L.grcuda_dmr_chain_seek_async(self.h, C.c_longlong(abs_row), _sp(stream)))

    def tell(self):
        return int(self.L.grcuda_dmr_chain_tell(self.h))

    def state_bytes(self):
        return int(self.L.grcuda_dmr_chain_state_bytes(self.h))

    def export_state(self, d_state, stream=None):
        _l.check(self.L.grcuda_dmr_chain_export_state(self.h, C.c_void_p(d_state.data_ptr()),
                                                      _sp(stream)))

    def import_state(self, d_state, stream=None):
        _l.check(self.L.grcuda_dmr_chain_import_state(self.h, C.c_void_p(d_state.data_ptr()),
                                                      _sp(stream)))

    def process_device(self, d_in, nrows, stream=None):
        """d_in: torch CUDA tensor (or raw pointer int) addressing history_rows() rows + nrows new rows.
        stream: a CUDA stream handle (e.g. torch.cuda.current_stream().cuda_stream) the front is ordered on; None =
        the chain's own (non-blocking) stream, which is NOT ordered after work queued on other streams: d_in must
        be complete before the call."""
        ptr = d_in if isinstance(d_in, int) else d_in.data_ptr()
        _l.check(self.L.grcuda_dmr_chain_process_device(self.h, C.c_void_p(ptr), int(nrows),
                                                        _sp(stream)))

    def join(self, stream=None):
        """Makes `stream` wait for the tails (M&M + slicer + correlator) queued so far on the chain's own stream."""
        _l.check(self.L.grcuda_dmr_chain_join(self.h, _sp(stream)))

    def process_front_device(self, d_in, nrows, stream=None):
        ptr = d_in if isinstance(d_in, int) else d_in.data_ptr()
        _l.check(self.L.grcuda_dmr_chain_process_front_device(self.h, C.c_void_p(ptr), int(nrows),
                                                              _sp(stream)))

    def process_tail_device(self, stream=None):
        _l.check(self.L.grcuda_dmr_chain_process_tail_device(self.h, _sp(stream)))

    def process_tail_mm_device(self, state_in=None, state_out=None, stream=None):
        """Clock recovery + slicer only; state_in / state_out: uint8 CUDA tensors (or None) of mm_state_bytes()."""
        _l.check(self.L.grcuda_dmr_chain_process_tail_mm_device(
            self.h, C.c_void_p(state_in.data_ptr() if state_in is not None else None),
            C.c_void_p(state_out.data_ptr() if state_out is not None else None), _sp(stream)))

    def process_tail_corr_device(self, state_in=None, state_out=None, stream=None):
        _l.check(self.L.grcuda_dmr_chain_process_tail_corr_device(
            self.h, C.c_void_p(state_in.data_ptr() if state_in is not None else None),
            C.c_void_p(state_out.data_ptr() if state_out is not None else None), _sp(stream)))

    def mm_state_bytes(self):
        return int(self.L.grcuda_dmr_chain_mm_state_bytes(self.h))

    def corr_state_bytes(self):
        return int(self.L.grcuda_dmr_chain_corr_state_bytes(self.h))

    STAGES = ("pfb_fir", "pfb_fft", "quad_demod", "rrc_fir", "mm_slicer", "map_unpack_corr", "carry_copies")

    def set_keep_channels(self, on):
        """False: the channelizer output is never written to HBM (fetch()["channels"] is absent): the discriminator runs
        inside the channelizer's FFT kernel.  Same symbols and hits.  NotImplementedError when the plan has no such kernel."""
        _l.check(self.L.grcuda_dmr_chain_set_keep_channels(self.h, int(on)))   # False/0, True/1, or 2 (fused kernel + stores)

    def keeps_channels(self):
        return int(self.L.grcuda_dmr_chain_keeps_channels(self.h))

    def set_tail_variant(self, variant):
        """Which build of the clock-recovery kernel the tail runs (0 = sized to co-reside with the front kernels)."""
        _l.check(self.L.grcuda_dmr_chain_set_tail_variant(self.h, int(variant)))

    def counters(self):
        """dict(clamped, overflow, hits_dropped): all zero unless the output has left the reference's (gr_cuda.h)."""
        a, b, c = C.c_longlong(), C.c_longlong(), C.c_longlong()
        _l.check(self.L.grcuda_dmr_chain_counters(self.h, C.byref(a), C.byref(b), C.byref(c)))
        return {"clamped": a.value, "overflow": b.value, "hits_dropped": c.value}

    def set_accumulate_hits(self, on):
        _l.check(self.L.grcuda_dmr_chain_set_accumulate_hits(self.h, int(bool(on))))

    def clear_hits(self, stream=None):
        _l.check(self.L.grcuda_dmr_chain_clear_hits(self.h, _sp(stream)))

    def max_hits(self):
        return int(self.L.grcuda_dmr_chain_max_hits(self.h))

    def set_split_correlator(self, on):
        _l.check(self.L.grcuda_dmr_chain_set_split_correlator(self.h, int(bool(on))))

    def set_profiling(self, on):
        _l.check(self.L.grcuda_dmr_chain_set_profiling(self.h, int(bool(on))))

    def profile_read(self):
        """{stage: (milliseconds, launches)} accumulated since the last read (CUDA events on the launch stream)."""
        ms = (C.c_float * 7)()
        ln = (C.c_int * 7)()
        _l.check(self.L.grcuda_dmr_chain_profile_read(self.h, ms, ln))
        return {n: (float(ms[i]), int(ln[i])) for i, n in enumerate(self.STAGES)}

    def process_host(self, rows, nrows):
        """rows: host array/pointer with history_rows() + nrows rows of M complex64 (pinned or pageable)."""
        if isinstance(rows, int):
            ptr = C.c_void_p(rows)
        else:
            rows = np.ascontiguousarray(rows, np.complex64)
            ptr = rows.ctypes.data_as(C.c_void_p)
        _l.check(self.L.grcuda_dmr_chain_process_host(self.h, ptr, int(nrows)))

    def result(self):
        r = _l.ChainResult()
        _l.check(self.L.grcuda_dmr_chain_result_get(self.h, C.byref(r)))
        return r

    HIT_DTYPE = np.dtype([("channel", np.int32), ("pad", np.int32), ("bit_index", np.int64)])   # == grcuda_hit

    def read_hits_array(self, max_hits=1 << 20):
        """Sync hits of the last block as a numpy record array (fields channel, bit_index) + the total count.
        One D2H copy into the array's memory: no per-hit Python work."""
        if getattr(self, "_hitbuf", None) is None or len(self._hitbuf) < max_hits:
            self._hitbuf = np.empty(max_hits, self.HIT_DTYPE)
        n = _l.check(self.L.grcuda_dmr_chain_read_hits(self.h, self._hitbuf.ctypes.data_as(C.c_void_p), int(max_hits)))
        return self._hitbuf[:min(n, max_hits)], n

    def read_hits(self, max_hits=1 << 20):
        """[(channel, absolute bit index), ...] of the last block, and the total count."""
        a, n = self.read_hits_array(max_hits)
        return list(zip(a["channel"].tolist(), a["bit_index"].tolist())), n

    # -- result readback helpers (tests / examples; use torch for big device-side consumers) --
    def _d2h(self, ptr, nbytes):
        out = np.empty(nbytes, np.uint8)
        _l.check(self.L.grcuda_memcpy_d2h(out.ctypes.data_as(C.c_void_p), C.c_void_p(ptr), C.c_size_t(nbytes), None))
        return out

    def fetch(self):
        """Copies the last block's results to numpy: dict(channels, soft, symbols, counts, bytes)."""
        self.L.grcuda_device_synchronize()
        r = self.result()
        M, ms = self.M, r.max_sym
        out = {
            "counts": self._d2h(r.d_sym_counts, M * 4).view(np.int32),
            "soft": self._d2h(r.d_soft, ms * M * 4).view(np.float32).reshape(ms, M),
            "symbols": self._d2h(r.d_symbols, ms * M).reshape(ms, M),
        }
        if r.d_channels:
            out["channels"] = self._d2h(r.d_channels, r.nrows * M * 8).view(np.complex64).reshape(r.nrows, M)
        if r.d_bytes:
            out["bytes"] = self._d2h(r.d_bytes, 2 * ms * M).reshape(2 * ms, M)
        return out
