"""digital_clock_recovery_mm_cc, oracle side only: the plain-C restatement of the complex
Mueller & Mueller loop (and of gri_mmse_fir_interpolator_cc) against the compiled reference class and the committed
fixture.  A feedback loop on float data: the comparison is bit for bit, against the reference built with the
generic-order gr_fir_ccf (the restatement's summation order).  The GPU block for this row is not written yet."""
import numpy as np
import pytest


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def test_oracle_fixture(orc, golden_next):
    fx = golden_next
    args = [float(v) for v in fx["mmcc_args"]]
    y, e, c = orc.mmcc_work(orc.mmcc_new(*args), fx["mmcc_x"])
    assert c == int(fx["mmcc_consumed_plain"]) and np.array_equal(bits(y), bits(fx["mmcc_y_plain"]))
    y, e, c = orc.mmcc_work(orc.mmcc_new(*args), fx["mmcc_x"], with_error=True)
    assert c == int(fx["mmcc_consumed_err"]) and np.array_equal(bits(y), bits(fx["mmcc_y_err"]))
    assert np.array_equal(bits(e), bits(fx["mmcc_err"]))
    assert np.abs(e).max() <= 4.0                                    # clipped to +-4 with the error output (:146)


def test_oracle_live_vs_reference_chunked(orc, ref):
    rng = np.random.default_rng(23)
    ref.set_fir_impl(0)
    try:
        for omega, gm, n in ((2.0, 0.05, 4000), (4.0, 0.1, 6000), (8.0, 0.175, 8000), (1.0, 0.01, 2000), (2.6041667, 0.175, 5000)):
            sps = int(np.ceil(omega))
            sym = (rng.integers(0, 2, n) * 2 - 1) + 1j * (rng.integers(0, 2, n) * 2 - 1)
            x = (np.repeat(sym, sps)[:n] + 0.1 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
            for we in (False, True):
                blk = ref.clock_recovery_mm_cc(omega, 0.25 * gm * gm, 0.5, gm, 0.005)
                st = orc.mmcc_new(omega, 0.25 * gm * gm, 0.5, gm, 0.005)
                assert blk.forecast(100) == orc.lib().orc_mmcc_forecast(__import__("ctypes").byref(st), 100)
                pos = 0
                while pos < n - 64:                                   # scheduler-style: limited output per call
                    yr, er, cr = ref.run_mm_cc(blk, x[pos:], noutput=257, with_error=we)
                    yo, eo, co = orc.mmcc_work(st, x[pos:], noutput=257, with_error=we)
                    assert cr == co and np.array_equal(bits(yr), bits(yo))
                    if we:
                        assert np.array_equal(bits(er), bits(eo))
                    if co == 0:
                        break
                    pos += co
    finally:
        ref.set_fir_impl(1)


def test_oracle_constructor_errors(orc, ref):
    with pytest.raises(IndexError):
        orc.mmcc_new(0.0, 0.1, 0.5, 0.1)
    with pytest.raises(IndexError):
        orc.mmcc_new(2.0, -0.1, 0.5, 0.1)
    with pytest.raises(IndexError):
        ref.clock_recovery_mm_cc(0.0, 0.1, 0.5, 0.1)
    with pytest.raises(IndexError):
        ref.clock_recovery_mm_cc(2.0, 0.1, 0.5, -0.1)
