"""gr_pfb_decimator_ccf: oracle against the compiled reference and the committed fixtures (CPU),
the CUDA path through the C ABI against both (GPU).  FIR-class output: bar 1e-4 of the output peak; measured <= 3e-7.
The reference de-spins with a decim-point FFT (FFTW, absent: float64 DFT in the oracle); one bin needs no FFT."""
import numpy as np
import pytest

from conftest import has_cuda

TOL = 1e-4
TAGS = ("m10", "m32", "m160")


def relerr(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30)) if a.size else 0.0


def crandn(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_fixture(orc, golden_next, tag):
    fx = golden_next
    M, ch = [int(v) for v in fx["pfbdec_%s_args" % tag]]
    y = orc.pfb_decimator_ccf(M, fx["pfbdec_%s_taps" % tag], ch, fx["pfbdec_%s_x" % tag])
    assert relerr(y, fx["pfbdec_%s_y" % tag]) < 1e-6


def test_oracle_live_vs_reference(orc, ref):
    rng = np.random.default_rng(11)
    for M, T, ch in ((2, 9, 1), (4, 8, 0), (7, 3, 2), (8, 5, 3), (20, 6, 19), (64, 4, 5)):
        taps = (rng.standard_normal(M * T - 1) * 0.1).astype(np.float32)
        x = crandn(rng, M * 200)
        blk = ref.pfb_decimator_ccf(M, taps, ch)
        assert blk.history == T
        want = ref.run_pfb_decimator(blk, x, M, chunk=61)
        assert relerr(orc.pfb_decimator_ccf(M, taps, ch, x), want) < 1e-6


def test_oracle_is_the_matching_channelizer_output(orc):
    """A polyphase decimator on channel c is channel c of the channelizer built from the same prototype, up to the
    different stream/filter pairing of the two blocks: a tone at c * fs / M comes out as a constant."""
    M, T, c = 8, 6, 3
    taps = np.hanning(M * T).astype(np.float32)
    taps /= taps.sum()
    n = np.arange(M * 400)
    x = np.exp(2j * np.pi * c * n / M).astype(np.complex64)
    y = orc.pfb_decimator_ccf(M, taps, c, x)[T:]
    assert np.abs(np.abs(y) - 1.0).max() < 1e-3 and np.abs(np.diff(y)).max() < 1e-4
    for other in (0, 2, 5):
        assert np.abs(orc.pfb_decimator_ccf(M, taps, other, x)[T:]).max() < 0.1


@pytest.fixture(scope="module")
def B():
    from grb200 import blocks
    return blocks


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_fixture_both_host_forms(B, golden_next, tag):
    fx = golden_next
    M, ch = [int(v) for v in fx["pfbdec_%s_args" % tag]]
    taps, x, want = fx["pfbdec_%s_taps" % tag], fx["pfbdec_%s_x" % tag], fx["pfbdec_%s_y" % tag]
    assert relerr(B.pfb_decimator_ccf(M, taps, ch).run(x, chunk=50), want) < TOL
    blk = B.pfb_decimator_ccf(M, taps, ch)
    T, n = blk.history(), len(x) // M
    streams = [np.concatenate([np.zeros(T - 1, np.complex64), x[s::M]]) for s in range(M)]
    assert blk.work(n, streams).size == 0                     # first work(): history may have changed
    assert relerr(blk.work(n, streams), want) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("M,T,ch,n", [(2, 9, 1, 5000), (4, 64, 3, 20000), (7, 3, 2, 1000), (20, 6, 19, 3000), (64, 4, 5, 2000),
                                      (160, 16, 159, 700), (1000, 5, 17, 64), (8000, 2, 4321, 40), (1, 5, 0, 100)])
def test_gpu_vs_oracle(B, orc, M, T, ch, n):
    rng = np.random.default_rng(M + T)
    taps = (rng.standard_normal(M * T - (1 if M * T > 1 else 0)) * 0.1).astype(np.float32)
    x = crandn(rng, M * n)
    got = B.pfb_decimator_ccf(M, taps, ch).run(x, chunk=n // 3 + 1)
    assert relerr(got, orc.pfb_decimator_ccf(M, taps, ch, x)) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_gpu_contract_and_set_taps(B, orc):
    rng = np.random.default_rng(2)
    M, ch = 8, 5
    t1 = (rng.standard_normal(M * 4) * 0.1).astype(np.float32)
    t2 = (rng.standard_normal(M * 9 - 3) * 0.1).astype(np.float32)
    blk = B.pfb_decimator_ccf(M, t1, ch)
    assert blk.history() == 4 and blk.taps_per_filter() == 4
    blk.set_taps(t2)
    assert blk.history() == 9                                  # set_history(taps_per_filter) in set_taps (:107)
    x = crandn(rng, M * 500)
    assert relerr(blk.run(x), orc.pfb_decimator_ccf(M, t2, ch, x)) < TOL
    with pytest.raises(ValueError):
        B.pfb_decimator_ccf(0, t1, 0)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_gpu_full_size_tone_property(B):
    """2 MS/s -> one 12.5 kHz channel of 160 (cfg3's shape) over 16 M samples, device resident: a tone on the channel
    comes out as a constant of the prototype's DC gain, a tone on another channel is rejected."""
    import torch
    from grb200 import firdes
    M, T, rows, c = 160, 16, 100000, 37
    taps = np.asarray(firdes.low_pass(1.0, 2e6, 5000.0, 2500.0), np.float32)[:M * T]
    blk = B.pfb_decimator_ccf(M, taps, c)
    n = torch.arange((rows + T - 1) * M, device="cuda", dtype=torch.float64)
    for chan, expect_pass in ((c, True), (c + 3, False)):
        ph = 2 * np.pi * chan * n / M
        x = torch.complex(torch.cos(ph).float(), torch.sin(ph).float()).reshape(rows + T - 1, M)
        y = torch.empty(rows, dtype=torch.complex64, device="cuda")
        b2 = B.pfb_decimator_ccf(M, taps, c)
        assert b2.work_device(rows, x, y) == 0
        assert b2.work_device(rows, x, y) == rows
        torch.cuda.synchronize()
        mag = y.abs()
        if expect_pass:
            assert abs(mag.mean().item() - float(np.sum(taps))) < 2e-3 and (mag.max() - mag.min()).item() < 1e-3
        else:
            assert mag.max().item() < 2e-3
    del blk
