"""CPU replay of the per-thread device functions (csrc/gr_math.cuh, csrc/fft_radix.cuh) against
the oracle: catches arithmetic / ordering mistakes before any GPU time is spent.  The GPU parity
tests (tests/test_gpu_*.py) remain the gate for the kernels themselves."""
import shutil

import numpy as np
import pytest

pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not available")


@pytest.fixture(scope="module")
def em():
    import emulharness
    emulharness.lib()
    return emulharness


PLANS = [(32, [8, 4]), (160, [16, 10]), (4096, [16, 16, 16]), (8000, [20, 20, 20]), (100, [10, 10]),
         (60, [4, 3, 5]), (1000, [10, 10, 10]), (400, [20, 20]), (2048, [16, 16, 8]), (30, [2, 3, 5]), (16, [16]),
         (20, [20]), (10, [10]), (8, [8]), (5, [5]), (3, [3])]


@pytest.mark.parametrize("N,radices", PLANS)
def test_stockham_passes_match_dft(em, N, radices):
    rng = np.random.default_rng(N)
    x = (rng.standard_normal(N) + 1j * rng.standard_normal(N)).astype(np.complex64)
    for dirn, ref in ((-1, np.fft.fft(x.astype(np.complex128))), (1, np.fft.ifft(x.astype(np.complex128)) * N)):
        y = em.fft(x, radices, dirn)
        err = np.max(np.abs(y - ref)) / np.max(np.abs(ref))
        assert err < 2e-6, (N, dirn, err)


def test_atan2_and_quad_demod_bit_exact(em, golden):
    fx = golden[1]
    assert np.array_equal(em.fast_atan2f(fx["atan_y"], fx["atan_x"]), fx["atan_out"])
    assert np.array_equal(em.quad_demod(float(fx["quad_gain"]), fx["quad_x"]), fx["quad_y"])


@pytest.mark.parametrize("ntaps", [1, 3, 4, 8, 13, 29, 64, 111])
def test_strided_fir_fff_both_orders(em, orc, ntaps):
    rng = np.random.default_rng(ntaps)
    taps = rng.standard_normal(ntaps).astype(np.float32)
    x = rng.standard_normal(300).astype(np.float32)
    for order in (orc.ORDER_GENERIC, orc.ORDER_SSE):
        assert np.array_equal(em.fir_fff(taps, x, order, nchan=5, chan=3), orc.fir_fff(taps, 1, x, order=order))


def test_mm_bit_exact(em, orc, golden):
    fx = golden[1]
    a = [float(v) for v in fx["mm_args"]]
    for order, nm in ((orc.ORDER_SSE, "sse"), (orc.ORDER_GENERIC, "generic")):
        y, c, st = em.mm(a, fx["mm_x"], order)
        assert c == int(fx["mm_consumed_" + nm]) and np.array_equal(y, fx["mm_y_" + nm])
    # unaligned absolute start exercises every SSE alignment phase
    x = fx["mm_x"]
    for abs0 in (1, 2, 3):
        y, c, _ = em.mm(a, x, orc.ORDER_SSE, abs0=abs0)
        y2, c2 = orc.mm_work(orc.mm_new(*a), x, order=orc.ORDER_SSE, abs0=abs0)
        assert c == c2 and np.array_equal(y, y2)


def test_slicer_and_correlator(em, orc, golden):
    fx = golden[1]
    assert np.array_equal(em.slice4(0.0, fx["slicer_x"]), fx["slicer_y_a0"])
    assert np.array_equal(em.slice4(0.01, fx["slicer_x"]), fx["slicer_y_a01"])
    from grb200 import synth
    assert np.array_equal(em.corr(synth.DMR_BS_DATA_SYNC_BITS, 2, fx["corr_bits"]), fx["corr_out_t2"])
