"""CPU-side checks of the drop-in boundary: libgr_cuda.so loads, exports every symbol declared in
include/gr_cuda.h, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import shutil

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200", "libgr_cuda.so")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(SO):
        if shutil.which("nvcc") is None:
            pytest.skip("libgr_cuda.so not built and nvcc not available")
        import __graft_entry__
        __graft_entry__.build()
    return ctypes.CDLL(SO)


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "gr_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(grcuda_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) > 80
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_no_cpu_fallback(lib):
    from conftest import has_cuda
    if has_cuda():
        pytest.skip("a GPU is present")
    lib.grcuda_fir_filter_ccf_create.restype = ctypes.c_void_p
    lib.grcuda_last_error.restype = ctypes.c_char_p
    taps = (ctypes.c_float * 4)(1, 2, 3, 4)
    assert lib.grcuda_device_count() == 0
    assert not lib.grcuda_fir_filter_ccf_create(1, taps, 4)
    assert lib.grcuda_last_error_code() == -3
    assert b"no CPU fallback" in lib.grcuda_last_error()
    from grb200 import blocks, lib as gl
    with pytest.raises(gl.GrCudaError):
        blocks.pfb_channelizer_ccf(8, [1.0] * 32)


def test_argument_errors_do_not_need_a_gpu(lib):
    """The reference constructors throw before touching any buffer; so do ours."""
    lib.grcuda_clock_recovery_mm_ff_create.restype = ctypes.c_void_p
    lib.grcuda_correlate_access_code_bb_create.restype = ctypes.c_void_p
    f = ctypes.c_float
    assert not lib.grcuda_clock_recovery_mm_ff_create(1, f(0.5), f(0.1), f(0.5), f(0.1), f(0.001), 1)
    assert lib.grcuda_last_error_code() == -2   # std::out_of_range
    assert not lib.grcuda_correlate_access_code_bb_create(1, b"1" * 65, 0)
    assert lib.grcuda_last_error_code() == -2


def test_argument_errors_of_the_8f_blocks(lib):
    """gr_pfb_arb_resampler_ccf / gr_pfb_decimator_ccf / gr_fft_filter_ccc: bad constructor arguments are refused with
    GRCUDA_EINVAL before any device is touched."""
    f = ctypes.c_float
    taps = (ctypes.c_float * 8)(*([0.125] * 8))
    for name in ("grcuda_pfb_arb_resampler_ccf_create", "grcuda_pfb_decimator_ccf_create", "grcuda_fft_filter_ccc_create"):
        getattr(lib, name).restype = ctypes.c_void_p
    assert not lib.grcuda_pfb_arb_resampler_ccf_create(f(1.5), taps, 1, 32, 1)      # create_diff_taps needs 2 taps
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_pfb_arb_resampler_ccf_create(f(0.0), taps, 8, 32, 1)      # rate must be > 0
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_pfb_arb_resampler_ccf_create(f(1.5), taps, 8, 0, 1)       # no filters
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_pfb_decimator_ccf_create(0, taps, 8, 0)
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_fft_filter_ccc_create(0, taps, 4)
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_fft_filter_ccc_create(1, taps, 0)
    assert lib.grcuda_last_error_code() == -1


def test_argument_errors_of_the_round2_blocks(lib):
    """gr_unpack_k_bits_bb, digital_clock_recovery_mm_cc, gr_framer_sink_1, gr_stream_to_streams, gr_map_bb: the reference's
    constructor exceptions (std::out_of_range for k == 0, omega <= 0, negative gains) and plain bad arguments come back
    as error codes before any device is touched."""
    f = ctypes.c_float
    for name in ("grcuda_unpack_k_bits_bb_create", "grcuda_clock_recovery_mm_cc_create", "grcuda_framer_sink_1_create",
                 "grcuda_stream_to_streams_create", "grcuda_vector_to_streams_create", "grcuda_map_bb_create"):
        getattr(lib, name).restype = ctypes.c_void_p
    lib.grcuda_stream_to_streams_create.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
    lib.grcuda_vector_to_streams_create.argtypes = [ctypes.c_size_t, ctypes.c_size_t]
    lib.grcuda_framer_sink_1_create.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t]
    assert not lib.grcuda_unpack_k_bits_bb_create(0)                                  # gr_unpack_k_bits_bb.cc:44-45
    assert lib.grcuda_last_error_code() == -2
    assert not lib.grcuda_clock_recovery_mm_cc_create(1, f(0.0), f(0.1), f(0.5), f(0.1), f(0.001))   # :65-66
    assert lib.grcuda_last_error_code() == -2
    assert not lib.grcuda_clock_recovery_mm_cc_create(1, f(2.0), f(-0.1), f(0.5), f(0.1), f(0.001))  # :67-68
    assert lib.grcuda_last_error_code() == -2
    assert not lib.grcuda_clock_recovery_mm_cc_create(0, f(2.0), f(0.1), f(0.5), f(0.1), f(0.001))
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_framer_sink_1_create(0, 16, 1024)
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_stream_to_streams_create(0, 4)
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_vector_to_streams_create(8, 0)
    assert lib.grcuda_last_error_code() == -1
    assert not lib.grcuda_map_bb_create(None, 4)
    assert lib.grcuda_last_error_code() == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                txt = open(os.path.join(dp, f)).read()
                for needle in ("import orc", "refharness", "liboracle", "libgrref", '#include "oracle', "#include <oracle"):
                    assert needle not in txt, (f, needle)
