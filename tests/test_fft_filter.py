"""gr_fft_filter_ccc: oracle restatement of the overlap-add filter against the compiled reference
and the committed fixtures (CPU); the CUDA path (direct-form FIR behind the same block contract) through the C ABI
against both (GPU).  FIR-class output: bar 1e-4 of the output peak; measured <= 3e-7.  Like the reference's own QA
(qa_fft_filter.py) the result is also compared with the plain time-domain convolution."""
import numpy as np
import pytest

from conftest import has_cuda

TOL = 1e-4
TAGS = ("d1_t33", "d3_t64", "d1_t200")


def relerr(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30)) if a.size else 0.0


def crandn(rng, n):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)


def conv(x, taps, decim, n):
    return np.convolve(x.astype(np.complex128), taps.astype(np.complex128))[:len(x)][::decim][:n].astype(np.complex64)


@pytest.mark.parametrize("tag", TAGS)
def test_oracle_fixture(orc, golden_next, tag):
    fx = golden_next
    dec, ns = [int(v) for v in fx["fftfilt_%s_args" % tag]]
    o = orc.FftFilter(dec, fx["fftfilt_%s_taps" % tag])
    assert o.nsamples == ns
    y = o.run(fx["fftfilt_x"], blocks_per_call=3)
    want = fx["fftfilt_%s_y" % tag]
    assert np.array_equal(y.view(np.uint32), want.view(np.uint32))        # same float64 DFT on both sides
    assert relerr(y, conv(fx["fftfilt_x"], fx["fftfilt_%s_taps" % tag], dec, len(y))) < 1e-6


def test_oracle_live_vs_reference_and_set_taps(orc, ref):
    rng = np.random.default_rng(13)
    for dec, nt, n in ((1, 1, 300), (1, 5, 2000), (2, 33, 6000), (4, 100, 9000), (3, 17, 4000), (1, 128, 3000)):
        taps = crandn(rng, nt) * np.float32(0.1)
        x = crandn(rng, n)
        blk = ref.fft_filter_ccc(dec, taps)
        o = orc.FftFilter(dec, taps)
        assert blk.output_multiple == o.nsamples and blk.history == 1
        assert relerr(o.run(x, 2), ref.run_fft_filter(blk, x, dec, blocks_per_call=3)) < 1e-6
    # set_taps: deferred, the next work() returns 0, the tail is cleared and the block size changes
    t1, t2 = crandn(rng, 20) * np.float32(0.1), crandn(rng, 70) * np.float32(0.1)
    blk = ref.fft_filter_ccc(1, t1)
    x = crandn(rng, 3000)
    ref.run_fft_filter(blk, x[:900], 1)
    ref.fft_filter_ccc_set_taps(blk, t2)
    out = np.zeros(64, np.complex64)
    raw, ptr, _ = ref.aligned_stream(x, 1, np.complex64)
    assert blk.general_work(45, [ptr], [45], out) == 0 and blk.output_multiple == 256 - 70 + 1
    o = orc.FftFilter(1, t1)
    o.run(x[:900])
    o.set_taps(t2)
    assert relerr(o.run(x), ref.run_fft_filter(blk, x, 1)) < 1e-6


@pytest.fixture(scope="module")
def B():
    from grb200 import blocks
    return blocks


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("tag", TAGS)
def test_gpu_fixture(B, golden_next, tag):
    fx = golden_next
    dec, ns = [int(v) for v in fx["fftfilt_%s_args" % tag]]
    blk = B.fft_filter_ccc(dec, fx["fftfilt_%s_taps" % tag])
    assert blk.output_multiple() == ns and blk.history() == 1
    for bpc in (None, 2):
        y = B.fft_filter_ccc(dec, fx["fftfilt_%s_taps" % tag]).run(fx["fftfilt_x"], bpc)
        assert relerr(y, fx["fftfilt_%s_y" % tag]) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("dec,nt,n", [(1, 1, 300), (1, 5, 2000), (2, 33, 60000), (4, 100, 90000), (3, 17, 4000), (1, 1000, 50000),
                                     (16, 257, 100000)])
def test_gpu_vs_oracle_and_time_domain(B, orc, dec, nt, n):
    rng = np.random.default_rng(nt)
    taps = crandn(rng, nt) * np.float32(0.1)
    x = crandn(rng, n)
    y = B.fft_filter_ccc(dec, taps).run(x, blocks_per_call=3)
    assert relerr(y, conv(x, taps, dec, len(y))) < TOL
    if n <= 10000:
        assert relerr(y, orc.FftFilter(dec, taps).run(x)) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
@pytest.mark.parametrize("dec,nt,n,path", [(1, 1, 300, 1), (1, 2, 500, 1), (1, 5, 2000, 1), (3, 17, 4000, 1), (2, 33, 60000, 1),
                                          (1, 129, 70000, 1), (4, 1000, 90000, 1), (1, 4096, 200000, 1), (5, 2049, 200000, 1),
                                          (1, 5, 2000, 2), (2, 100, 30000, 2), (1, 4097, 300000, 2), (3, 20000, 600000, 2),
                                          (1, 4097, 300000, -1), (1, 600, 50000, -1), (1, 20, 5000, -1)])
def test_gpu_frequency_domain_paths(B, orc, dec, nt, n, path):
    """The long filters that are the block's reason to exist, on both frequency-domain paths (and the automatic choice),
    against the time-domain convolution in float64 and (small cases) the oracle's overlap-add."""
    rng = np.random.default_rng(nt + 7 * dec)
    taps = crandn(rng, nt) * np.float32(1.0 / np.sqrt(nt))
    x = crandn(rng, n)
    blk = B.fft_filter_ccc(dec, taps)
    if path >= 0:
        blk.set_path(path)
        assert blk.work(blk.output_multiple(), x).size == 0       # deferred like set_taps
        assert blk.path() == path
    else:
        assert blk.path() == (0 if nt <= 32 else (1 if nt <= 4096 else 2))
    y = blk.run(x, blocks_per_call=3)
    ns = blk.output_multiple()
    assert len(y) == (n // dec) // ns * ns and len(y) > 0
    from scipy.signal import fftconvolve
    want = fftconvolve(x.astype(np.complex128), taps.astype(np.complex128))[: len(y) * dec: dec]
    assert relerr(y, want) < 5e-6
    if n <= 10000:
        assert relerr(y, orc.FftFilter(dec, taps).run(x)) < TOL


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_gpu_contract_set_taps_and_errors(B, orc):
    rng = np.random.default_rng(4)
    t1, t2 = crandn(rng, 20) * np.float32(0.1), crandn(rng, 70) * np.float32(0.1)
    x = crandn(rng, 3000)
    blk = B.fft_filter_ccc(1, t1)
    assert blk.output_multiple() == 64 - 20 + 1
    blk.run(x[:900])
    blk.set_taps(t2)
    assert blk.work(45, x[:45]).size == 0                      # "output multiple may have changed" (:85-90)
    assert blk.output_multiple() == 256 - 70 + 1
    o = orc.FftFilter(1, t2)
    assert relerr(blk.run(x), o.run(x)) < TOL                  # the carried state was cleared by set_taps
    with pytest.raises(ValueError):
        blk.work(100, x[:100])                                 # not a multiple of output_multiple (:92 assert)
    with pytest.raises(ValueError):
        B.fft_filter_ccc(0, t1)


@pytest.mark.gpu
@pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")
def test_gpu_device_form_full_size_linearity(B):
    """16 M samples, 129 complex taps, device resident: linearity and agreement of chunked and one-shot runs (the carried
    history makes the chunking invisible)."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(9)
    rng = np.random.default_rng(9)
    taps = crandn(rng, 129) * np.float32(0.05)
    ns = B.fft_filter_ccc(1, taps).output_multiple()
    n = ns * 125000
    x = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda"))
    y = torch.view_as_complex(torch.randn((n, 2), generator=g, device="cuda"))

    def run(v, pieces):
        blk = B.fft_filter_ccc(1, taps)
        out = torch.empty(n, dtype=torch.complex64, device="cuda")
        step = (n // ns // pieces) * ns
        done = 0
        while done < n:
            m = min(step, n - done)
            assert blk.work_device(m, v[done:], out[done:]) == m
            done += m
        torch.cuda.synchronize()
        return out

    rx = run(x, 1)
    r7 = run(x, 7)
    assert (rx - r7).abs().max().item() <= 1e-6 * rx.abs().max().item()
    rz = run(2.0 * x + y, 3)
    err = (rz - (2.0 * rx + run(y, 1))).abs().max().item() / rz.abs().max().item()
    assert err < 1e-5, err
