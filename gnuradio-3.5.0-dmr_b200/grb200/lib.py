"""ctypes loader for libgr_cuda.so (the C ABI in include/gr_cuda.h).

There is no CPU fallback anywhere in this package: if the shared library is missing, or there is
no CUDA device, every block constructor raises.  The oracle (oracle/) is never imported from here.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# GRCUDA_LIB: another build of the SAME library (instrumented lab builds of tools/, see csrc/Makefile); never a fallback
LIB_PATH = os.environ.get("GRCUDA_LIB") or os.path.join(_PKG, "libgr_cuda.so")

OK, EINVAL, ERANGE, ECUDA, ENOMEM, EUNSUPPORTED = 0, -1, -2, -3, -4, -5
ORDER_GENERIC, ORDER_SSE = 0, 1

_lib = None


class GrCudaError(RuntimeError):
    """std::runtime_error in the C++ wrappers (CUDA failure / no device)."""


class Hit(C.Structure):
    _fields_ = [("channel", C.c_int), ("pad", C.c_int), ("bit_index", C.c_longlong)]


class FramerMsg(C.Structure):
    _fields_ = [("channel", C.c_int), ("whitener_offset", C.c_int), ("length", C.c_int), ("seq", C.c_int),
                ("payload_offset", C.c_longlong), ("end_index", C.c_longlong)]


class ChainParams(C.Structure):
    _fields_ = [
        ("numchans", C.c_uint), ("pfb_taps", C.POINTER(C.c_float)), ("pfb_ntaps", C.c_int),
        ("quad_gain", C.c_float), ("rrc_taps", C.POINTER(C.c_float)), ("rrc_ntaps", C.c_int),
        ("omega", C.c_float), ("gain_omega", C.c_float), ("mu", C.c_float), ("gain_mu", C.c_float),
        ("omega_relative_limit", C.c_float), ("slicer_alpha", C.c_float),
        ("symbol_map", C.POINTER(C.c_int)), ("symbol_map_len", C.c_int),
        ("access_code", C.c_char_p), ("threshold", C.c_int), ("order", C.c_int),
        ("max_rows_per_block", C.c_int), ("keep_bytes", C.c_int),
    ]


class ChainResult(C.Structure):
    _fields_ = [
        ("d_channels", C.c_void_p), ("d_soft", C.c_void_p), ("d_symbols", C.c_void_p), ("d_sym_counts", C.c_void_p),
        ("d_bytes", C.c_void_p), ("d_hits", C.c_void_p), ("d_nhits", C.c_void_p), ("max_sym", C.c_int),
        ("nrows", C.c_int),
    ]


_PTR_RETURNING = [
    "grcuda_fir_filter_ccf_create", "grcuda_fir_filter_fff_create", "grcuda_freq_xlating_fir_filter_ccf_create",
    "grcuda_pfb_channelizer_ccf_create", "grcuda_fft_vcc_create", "grcuda_quadrature_demod_cf_create",
    "grcuda_clock_recovery_mm_ff_create", "grcuda_pager_slicer_fb_create", "grcuda_binary_slicer_fb_create",
    "grcuda_correlate_access_code_bb_create", "grcuda_dmr_chain_create", "grcuda_malloc_device",
    "grcuda_pfb_arb_resampler_ccf_create", "grcuda_pfb_decimator_ccf_create", "grcuda_fft_filter_ccc_create",
    "grcuda_malloc_pinned", "grcuda_map_bb_create", "grcuda_unpack_k_bits_bb_create", "grcuda_stream_to_streams_create",
    "grcuda_vector_to_streams_create", "grcuda_framer_sink_1_create", "grcuda_clock_recovery_mm_cc_create",
]


def load():
    """Returns the loaded library; raises if libgr_cuda.so has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrCudaError("%s not found: build it with __graft_entry__.build() / make -C csrc; "
                          "grb200 has no CPU fallback" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for n in _PTR_RETURNING:
        getattr(L, n).restype = C.c_void_p
    L.grcuda_last_error.restype = C.c_char_p
    L.grcuda_version.restype = C.c_char_p
    L.grcuda_kernel_launch_count.restype = C.c_ulonglong
    L.grcuda_pfb_channelizer_ccf_relative_rate.restype = C.c_double
    L.grcuda_quadrature_demod_cf_gain.restype = C.c_float
    L.grcuda_pager_slicer_fb_dc_offset.restype = C.c_float
    L.grcuda_dmr_chain_state_bytes.restype = C.c_size_t
    L.grcuda_dmr_chain_mm_state_bytes.restype = C.c_size_t
    L.grcuda_dmr_chain_corr_state_bytes.restype = C.c_size_t
    L.grcuda_dmr_chain_tell.restype = C.c_longlong
    L.grcuda_fir_filter_ccf_history.restype = C.c_uint
    L.grcuda_fir_filter_fff_history.restype = C.c_uint
    L.grcuda_freq_xlating_fir_filter_ccf_history.restype = C.c_uint
    L.grcuda_pfb_channelizer_ccf_history.restype = C.c_uint
    L.grcuda_pfb_arb_resampler_ccf_history.restype = C.c_uint
    L.grcuda_pfb_decimator_ccf_history.restype = C.c_uint
    L.grcuda_pfb_arb_resampler_ccf_relative_rate.restype = C.c_double
    L.grcuda_unpack_k_bits_bb_interpolation.restype = C.c_uint
    L.grcuda_stream_to_streams_create.argtypes = [C.c_size_t, C.c_size_t]
    L.grcuda_vector_to_streams_create.argtypes = [C.c_size_t, C.c_size_t]
    L.grcuda_framer_sink_1_create.argtypes = [C.c_int, C.c_int, C.c_size_t]
    _lib = L
    return L


def check(rc):
    """Negative return codes -> the exception type the reference block would throw."""
    if rc is None or rc >= 0:
        return rc
    msg = load().grcuda_last_error().decode()
    if rc == EINVAL:
        raise ValueError(msg)          # std::invalid_argument
    if rc == ERANGE:
        raise IndexError(msg)          # std::out_of_range
    if rc == ENOMEM:
        raise MemoryError(msg)
    if rc == EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise GrCudaError(msg)


def check_handle(h):
    if h:
        return C.c_void_p(h)
    L = load()
    code = L.grcuda_last_error_code()
    check(code if code < 0 else ECUDA)


def launches():
    return int(load().grcuda_kernel_launch_count())
