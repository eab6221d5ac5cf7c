"""optfir.low_pass / high_pass over gr.remez and window.blackmanharris -- the host-side design code
behind the DEFAULT taps of blks2.pfb_channelizer_ccf (blks2impl/pfb_channelizer.py:40-59).

Pinned (a) to tests/golden/ref_fixtures_optfir.npz (taps of the reference's gr_remez.cc compiled in place and of the
reference's window.py source executed as it stands; tests/golden/make_golden.py optfir) and (b) to the live
oracle/_ref library when it is there.  The design runs in float64; blocks take float32 taps, so the bar is: equal after
the float32 cast the block constructors apply, and <= 1e-12 of the peak in float64 (numpy evaluates cos() of the dense
grid with its own vector libm)."""
import numpy as np
import pytest

from grb200 import optfir, window, firdes

import os
FX = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_fixtures_optfir.npz"))


def design(name):
    kind, gain, fs, f1, f2, rip, att = FX["spec_" + name]
    f = optfir.low_pass if kind == 0 else optfir.high_pass
    return f(gain, fs, f1, f2, rip, att)


@pytest.mark.parametrize("name", ["lp8", "lp32", "lp48k", "hp"])
def test_remez_fixture(name):
    h, r = design(name), FX["remez_" + name]
    assert len(h) == len(r)
    assert np.max(np.abs(h - r)) <= 1e-12 * np.max(np.abs(r))
    assert np.array_equal(h.astype(np.float32), r.astype(np.float32))


def test_remez_live_reference(ref):
    # odd-symmetry designs start at 0 here: with bands[0] >= delf the reference writes one element past its malloc'd
    # grid (gr_remez.cc:640-647 shortens gridsize after counting) and glibc aborts the process
    for args in [(20, [0, 0.3, 0.4, 1.0], [1, 1, 0, 0], [1, 10]), (33, [0, 0.2, 0.3, 0.6, 0.7, 1], [0, 0, 1, 1, 0, 0], [5, 1, 5]),
                 (41, [0, 0.95], [1, 1], [], "hilbert"), (30, [0, 0.8], [0, 0.8 * np.pi / 2], [], "differentiator")]:
        h = optfir.remez(*args)
        r = ref.remez(*args)
        assert len(h) == len(r)
        assert np.max(np.abs(h - r)) <= 1e-11 * max(np.max(np.abs(r)), 1e-30), args


def test_remez_argument_errors():
    for bad in [(2, [0, 0.4, 0.6, 1], [1, 1, 0, 0]), (20, [0, 0.4, 0.6], [1, 1, 0]), (20, [0, 0.6, 0.4, 1], [1, 1, 0, 0]),
                (20, [0, 0.4, 0.6, 1.5], [1, 1, 0, 0]), (20, [0, 0.4, 0.6, 1], [1, 0]), (20, [0, 0.4, 0.6, 1], [1, 1, 0, 0], [1]),
                (20, [0, 0.4, 0.6, 1], [1, 1, 0, 0], [], "lowpass"), (20, [0, 0.4, 0.6, 1], [1, 1, 0, 0], [], "bandpass", 8)]:
        with pytest.raises(RuntimeError):
            optfir.remez(*bad)


def test_remezord_and_devs():
    # optfir.py:283-316 on the lp8 spec: the order estimate the reference prints in its own doc example
    n, fo, ao, w = optfir.remezord([0.4, 0.6], (1, 0), [optfir.passband_ripple_to_dev(0.1), optfir.stopband_atten_to_dev(60)], 8)
    assert n + 2 + 1 == len(FX["remez_lp8"])
    assert fo == [0, 0.1, 0.15, 1] and ao == [1, 1, 0, 0]
    assert w[0] == pytest.approx(1.0) or w[1] == pytest.approx(1.0)


@pytest.mark.parametrize("n", [64, 4096])
def test_blackmanharris_fixture(n):
    w = np.array(window.blackmanharris(n))
    assert np.array_equal(w, FX["blackmanharris_%d" % n])
    # and it is NOT gr_firdes::window(WIN_BLACKMAN_hARRIS): the reference carries two slightly different ones
    assert not np.allclose(w, firdes.window(firdes.WIN_BLACKMAN_hARRIS, n, 0), atol=1e-6)


def test_pfb_channelizer_default_taps_shape():
    """blks2.pfb_channelizer_ccf(numchans) with taps=None: the prototype is a low-pass at 0.4 of the channel spacing and
    >= 100 dB down from 0.6 on; the block then splits it over numchans branches."""
    M = 20
    taps = optfir.pfb_channelizer_default_taps(M)
    assert taps.dtype == np.float64 and len(taps) > 8 * M
    H = np.abs(np.fft.rfft(taps, 1 << 16))
    f = np.arange(len(H)) / (1 << 16) * M          # in channel spacings
    assert np.all(np.abs(20 * np.log10(H[f <= 0.4])) < 0.2)
    assert np.max(20 * np.log10(H[f >= 0.6] + 1e-300)) < -95.0


@pytest.mark.gpu
def test_default_taps_feed_the_gpu_channelizer(orc):
    """An unchanged flowgraph script: blks2.pfb_channelizer_ccf(numchans) designs its own taps, the GPU block takes
    them, the result equals the oracle channelizer on those taps."""
    from grb200 import blocks
    M = 20
    taps = optfir.pfb_channelizer_default_taps(M).astype(np.float32)
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(M * 600) + 1j * rng.standard_normal(M * 600)).astype(np.complex64)
    rows = 600
    want, wc = orc.pfb_channelizer_ccf(M, taps, x)
    blk = blocks.pfb_channelizer_ccf(M, taps)
    T = blk.taps_per_filter()
    inter = np.concatenate([np.zeros((T, M), np.complex64), x.reshape(rows, M)])
    assert len(blk.general_work_interleaved(rows, inter)[0]) == 0      # first call after set_taps returns 0
    got, c = blk.general_work_interleaved(rows, inter)
    assert c == wc == rows and got.shape == want.shape
    assert np.max(np.abs(got - want)) / np.max(np.abs(want)) < 5e-6


@pytest.mark.gpu
def test_blks2_pfb_channelizer_hier_block(orc):
    """blks2.pfb_channelizer_ccf(numchans) as a script writes it (no taps): one interleaved stream in, numchans streams
    out, equal to the oracle channelizer on the taps the reference's design rule gives."""
    from grb200 import blks2
    M = 16
    rng = np.random.default_rng(8)
    x = (rng.standard_normal(M * 400) + 1j * rng.standard_normal(M * 400)).astype(np.complex64)
    blk = blks2.pfb_channelizer_ccf(M)
    outs = blk.run(x)
    assert len(outs) == M
    want, _ = orc.pfb_channelizer_ccf(M, blk.taps().astype(np.float32), x)
    got = np.stack(outs, axis=1)
    n = min(len(got), len(want))
    assert n >= 390 and np.max(np.abs(got[:n] - want[:n])) / np.max(np.abs(want)) < 5e-6
    # a tone in channel 3 comes out of stream 3
    t = np.exp(2j * np.pi * 3 / M * np.arange(M * 200)).astype(np.complex64)
    p = np.array([np.abs(s[-50:]).mean() for s in blk.run(t)])
    assert np.argmax(p) == 3 and p[3] > 100 * np.delete(p, 3).max()


def test_optfir_band_pass_and_high_pass_live(ref):
    """optfir.band_pass / high_pass (optfir.py:75-88, 142-158): remezord's order estimate and band vector handed to the
    compiled gr_remez give the same taps as our whole design chain."""
    for args in [(1.0, 48000.0, 2000.0, 3000.0, 9000.0, 10500.0, 0.2, 60.0)]:
        h = optfir.band_pass(*args)
        pd, sd = optfir.passband_ripple_to_dev(args[6]), optfir.stopband_atten_to_dev(args[7])
        n, fo, ao, w = optfir.remezord(list(args[2:6]), (0, args[0], 0), [sd, pd, sd], args[1])
        r = ref.remez(n + 2, fo, ao, w)
        assert len(h) == len(r) == n + 3 and np.max(np.abs(h - r)) <= 1e-11 * np.max(np.abs(r))
        H = np.abs(np.fft.rfft(h, 1 << 14))
        f = np.arange(len(H)) / (1 << 14) * args[1]
        assert np.all(np.abs(20 * np.log10(H[(f >= 3000) & (f <= 9000)])) < 0.5)
        assert np.max(20 * np.log10(H[(f <= 2000) | (f >= 10500)] + 1e-300)) < -55
    h = optfir.high_pass(1.0, 8000.0, 1000.0, 1500.0, 0.1, 60.0)
    assert len(h) % 2 == 1                                    # optfir.high_pass forces an odd number of taps (:150-152)
