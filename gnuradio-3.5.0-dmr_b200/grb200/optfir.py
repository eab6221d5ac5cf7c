"""Equiripple FIR design on the host: gnuradio.optfir (gnuradio-core/src/python/gnuradio/optfir.py:45-62,115-360) over
gr.remez (gnuradio-core/src/lib/general/gr_remez.cc:98-877, the Parks-McClellan exchange in Janovetz's formulation).

These are the DEFAULT taps of blks2.pfb_channelizer_ccf (blks2impl/pfb_channelizer.py:40-59:
optfir.low_pass(1, numchans, 0.4, 0.6, ripple, atten=100)), so an unchanged flowgraph script needs them to build the
constructor argument of the GPU block.  Host-side design code, float64 like the reference; the dense-grid error
evaluation (the O(grid x extrema) part) is vectorised with numpy, everything that decides the exchange (grid, initial
guess, barycentric weights with the reference's skip pattern, the extremum search and its deletion rules, the stopping
test, the frequency sampling) follows the reference statement by statement.  tests/test_firdes_optfir.py pins it to the
compiled reference (oracle/_ref) and to fixtures.
"""
import math

import numpy as np

BANDPASS, DIFFERENTIATOR, HILBERT = 1, 2, 3
NEGATIVE, POSITIVE = 0, 1
PI = 3.14159265358979323846
PI2 = 2 * PI
GRIDDENSITY = 16
MAXITERATIONS = 40


def _dense_grid(r, numtaps, numband, bands, des, weight, gridsize, symmetry, griddensity):
    """CreateDenseGrid (gr_remez.cc:98-149)."""
    # For odd symmetry the reference shortens gridsize by one AFTER summing the per-band counts, yet fills all of them
    # (one element past its allocation when bands[0] >= delf); the slack keeps that write and the slice drops it.
    slack = numband + 2
    grid = np.zeros(gridsize + slack)
    D = np.zeros(gridsize + slack)
    W = np.zeros(gridsize + slack)
    delf = 0.5 / (griddensity * r)
    grid0 = delf if (symmetry == NEGATIVE and delf > bands[0]) else bands[0]
    j = 0
    for band in range(numband):
        lowf = grid0 if band == 0 else bands[2 * band]
        highf = bands[2 * band + 1]
        k = int((highf - lowf) / delf + 0.5)
        for i in range(k):
            D[j] = des[2 * band] + i * (des[2 * band + 1] - des[2 * band]) / (k - 1)
            W[j] = weight[band]
            grid[j] = lowf
            lowf += delf
            j += 1
        grid[j - 1] = highf
    if symmetry == NEGATIVE and grid[gridsize - 1] > (0.5 - delf) and numtaps % 2:
        grid[gridsize - 1] = 0.5 - delf
    return grid[:gridsize], D[:gridsize], W[:gridsize]


def _calc_parms(r, ext, grid, D, W):
    """CalcParms (gr_remez.cc:191-247): Lagrange weights ad, abscissae x, ordinates y."""
    x = np.cos(PI2 * grid[ext])
    ld = (r - 1) // 15 + 1  # "skips around to avoid round errors": the product is taken in ld interleaved passes
    order = np.concatenate([np.arange(j, r + 1, ld) for j in range(ld)])
    ad = np.empty(r + 1)
    for i in range(r + 1):
        ks = order[order != i]
        denom = 1.0
        for v in 2.0 * (x[i] - x[ks]):      # sequential product in the reference's order
            denom *= v
        if abs(denom) < 0.00001:
            denom = 0.00001
        ad[i] = 1.0 / denom
    sign = np.where(np.arange(r + 1) % 2 == 0, 1.0, -1.0)
    numer = denom = 0.0
    for i in range(r + 1):
        numer += ad[i] * D[ext[i]]
        denom += sign[i] * ad[i] / W[ext[i]]
    delta = numer / denom
    y = D[ext] - sign * delta / W[ext]
    return ad, x, y


def _compute_a(freqs, ad, x, y):
    """ComputeA (gr_remez.cc:269-290) on an array of frequencies: barycentric interpolation, with the reference's
    short-cut when a frequency coincides with an interpolation point."""
    out = np.empty(len(freqs))
    step = max(1, 4_000_000 // len(x))
    for a in range(0, len(freqs), step):
        xc = np.cos(PI2 * freqs[a:a + step])
        c = xc[:, None] - x[None, :]
        hit = np.abs(c) < 1.0e-7
        with np.errstate(divide="ignore", invalid="ignore"):
            cc = ad[None, :] / c
            # sequential sums in the reference's order (numpy's reductions are pairwise; for a few thousand taps the
            # barycentric sums are ill conditioned enough for the summation order to decide the exchange)
            val = np.cumsum(cc * y[None, :], axis=1)[:, -1] / np.cumsum(cc, axis=1)[:, -1]
        if hit.any():
            rows = np.nonzero(hit.any(1))[0]
            for rr in rows:   # the loop breaks at the FIRST coinciding point
                val[rr] = y[np.nonzero(hit[rr])[0][0]]
        out[a:a + step] = val
    return out


def _search(r, ext, gridsize, E):
    """Search (gr_remez.cc:357-470).  Returns 0, -2 (insufficient extremals) or -3 (too many)."""
    found = []
    if (E[0] > 0.0 and E[0] > E[1]) or (E[0] < 0.0 and E[0] < E[1]):
        found.append(0)
    e0, em, ep = E[1:-1], E[:-2], E[2:]
    mid = np.nonzero(((e0 >= em) & (e0 > ep) & (e0 > 0.0)) | ((e0 <= em) & (e0 < ep) & (e0 < 0.0)))[0] + 1
    if len(found) + len(mid) > 2 * r:
        return -3
    found.extend(int(i) for i in mid)
    j = gridsize - 1
    if (E[j] > 0.0 and E[j] > E[j - 1]) or (E[j] < 0.0 and E[j] < E[j - 1]):
        if len(found) >= 2 * r:
            return -3
        found.append(j)
    if len(found) < r + 1:
        return -2
    extra = len(found) - (r + 1)
    while extra > 0:
        k = len(found)
        up = E[found[0]] > 0.0
        l = 0
        alt = True
        for j in range(1, k):
            if abs(E[found[j]]) < abs(E[found[l]]):
                l = j
            if up and E[found[j]] < 0.0:
                up = False
            elif (not up) and E[found[j]] > 0.0:
                up = True
            else:
                alt = False
                break
        if alt and extra == 1:
            l = k - 1 if abs(E[found[k - 1]]) < abs(E[found[0]]) else 0
        del found[l]
        extra -= 1
    ext[:] = found[:r + 1]
    return 0


def _freq_sample(N, A, symm):
    """FreqSample (gr_remez.cc:494-560)."""
    h = np.empty(N)
    M = (N - 1.0) / 2.0
    n = np.arange(N)
    xx = PI2 * (n - M) / N
    if symm == POSITIVE:
        kmax = int(M) if N % 2 else N // 2 - 1
        k = np.arange(1, kmax + 1)
        val = A[0] + 2.0 * (A[k][None, :] * np.cos(xx[:, None] * k[None, :])).sum(1)
    else:
        if N % 2:
            k = np.arange(1, int(M) + 1)
            val = 2.0 * (A[k][None, :] * np.sin(xx[:, None] * k[None, :])).sum(1)
        else:
            k = np.arange(1, N // 2)
            val = A[N // 2] * np.sin(PI * (n - M)) + 2.0 * (A[k][None, :] * np.sin(xx[:, None] * k[None, :])).sum(1)
    h[:] = val / N
    return h


def _remez(numtaps, numband, bands, des, weight, ftype, griddensity):
    """remez (gr_remez.cc:612-780).  Returns (taps, err) with err 0 / -1 / -2 / -3 like the reference."""
    symmetry = POSITIVE if ftype == BANDPASS else NEGATIVE
    r = numtaps // 2
    if numtaps % 2 and symmetry == POSITIVE:
        r += 1
    gridsize = 0
    for i in range(numband):
        gridsize += int(2 * r * griddensity * (bands[2 * i + 1] - bands[2 * i]) + .5)
    if symmetry == NEGATIVE:
        gridsize -= 1
    grid, D, W = _dense_grid(r, numtaps, numband, bands, des, weight, gridsize, symmetry, griddensity)
    ext = [i * (gridsize - 1) // r for i in range(r + 1)]      # InitialGuess
    if ftype == DIFFERENTIATOR:
        m = D > 0.0001
        W[m] = W[m] / grid[m]
    if symmetry == POSITIVE:
        if numtaps % 2 == 0:
            c = np.cos(PI * grid)
            D /= c
            W *= c
    else:
        c = np.sin(PI2 * grid) if numtaps % 2 else np.sin(PI * grid)
        D /= c
        W *= c
    it = 0
    for it in range(MAXITERATIONS):
        ad, x, y = _calc_parms(r, np.asarray(ext), grid, D, W)
        E = W * (D - _compute_a(grid, ad, x, y))
        err = _search(r, ext, gridsize, E)
        if err:
            return None, err
        ee = np.abs(E[np.asarray(ext)])
        if (ee.max() - ee.min()) / ee.max() < 0.0001:      # isDone
            break
    else:
        it = MAXITERATIONS
    ad, x, y = _calc_parms(r, np.asarray(ext), grid, D, W)
    i = np.arange(numtaps // 2 + 1)
    if symmetry == POSITIVE:
        c = np.ones(len(i)) if numtaps % 2 else np.cos(PI * i / numtaps)
    else:
        c = np.sin(PI2 * i / numtaps) if numtaps % 2 else np.sin(PI * i / numtaps)
    taps = _compute_a(i / numtaps, ad, x, y) * c
    h = _freq_sample(numtaps, taps, symmetry)
    return h, (0 if it < MAXITERATIONS else -1)


def remez(order, bands, ampl, error_weight=(), filter_type="bandpass", grid_density=16):
    """gr.remez (gr_remez.cc:792-877): order + 1 taps; bands in [0, 1] (1 = Nyquist); RuntimeError like the reference."""
    numtaps = order + 1
    if numtaps < 4:
        raise RuntimeError("gr_remez: number of taps must be >= 3")
    bands = [float(b) for b in bands]
    numbands = len(bands) // 2
    if numbands < 1 or len(bands) % 2 == 1:
        raise RuntimeError("gr_remez: must have an even number of band edges")
    for i in range(1, len(bands)):
        if bands[i] < bands[i - 1]:
            raise RuntimeError("gr_remez: band edges must be nondecreasing")
    if bands[0] < 0 or bands[-1] > 1:
        raise RuntimeError("gr_remez: band edges must be in the range [0,1]")
    b2 = [b / 2 for b in bands]
    if len(ampl) != len(bands):
        raise RuntimeError("gr_remez: must have one response magnitude for each band edge")
    weight = [1.0] * numbands
    if len(error_weight) != 0:
        if len(error_weight) != numbands:
            raise RuntimeError("gr_remez: need one weight for each band [=length(band)/2]")
        weight = [float(w) for w in error_weight]
    itype = {"bandpass": BANDPASS, "differentiator": DIFFERENTIATOR, "hilbert": HILBERT}.get(filter_type)
    if itype is None:
        raise RuntimeError("gr_remez: unknown ftype '%s'" % filter_type)
    if grid_density < 16:
        raise RuntimeError("gr_remez: grid_density is too low; must be >= 16")
    h, err = _remez(numtaps, numbands, b2, [float(a) for a in ampl], weight, itype, grid_density)
    if err == -1:
        raise RuntimeError("gr_remez: failed to converge")
    if err == -2:
        raise RuntimeError("gr_remez: insufficient extremals -- cannot continue")
    if err == -3:
        raise RuntimeError("gr_remez: too many extremals -- cannot continue")
    return h


# ---- optfir.py ---------------------------------------------------------------------------------------------------
def stopband_atten_to_dev(atten_db):
    return 10 ** (-atten_db / 20.0)


def passband_ripple_to_dev(ripple_db):
    return (10 ** (ripple_db / 20.0) - 1) / (10 ** (ripple_db / 20.0) + 1)


def lporder(freq1, freq2, delta_p, delta_s):
    """Herrmann et al. length estimate (optfir.py:283-316)."""
    df = abs(freq2 - freq1)
    ddp = math.log10(delta_p)
    dds = math.log10(delta_s)
    a1, a2, a3, a4, a5, a6 = 5.309e-3, 7.114e-2, -4.761e-1, -2.66e-3, -5.941e-1, -4.278e-1
    b1, b2 = 11.01217, 0.5124401
    t1 = a1 * ddp * ddp
    t2 = a2 * ddp
    t3 = a4 * ddp * ddp
    t4 = a5 * ddp
    dinf = ((t1 + t2 + a3) * dds) + (t3 + t4 + a6)
    ff = b1 + b2 * (ddp - dds)
    return dinf / df - ff * df + 1


def remezord(fcuts, mags, devs, fsamp=2):
    """optfir.py:170-279."""
    fcuts = [float(f) / fsamp for f in fcuts]
    mags = list(mags)
    devs = list(devs)
    nbands = len(mags)
    if len(mags) != len(devs):
        raise ValueError("Length of mags and devs must be equal")
    if len(fcuts) != 2 * (nbands - 1):
        raise ValueError("Length of f must be 2 * len (mags) - 2")
    for i in range(len(mags)):
        if mags[i] != 0:
            devs[i] = devs[i] / mags[i]
    f1 = fcuts[0::2]
    f2 = fcuts[1::2]
    n = 0
    min_delta = 2
    for i in range(len(f1)):
        if f2[i] - f1[i] < min_delta:
            n = i
            min_delta = f2[i] - f1[i]
    if nbands == 2:
        l = lporder(f1[n], f2[n], devs[0], devs[1])
    else:
        l = 0
        for i in range(1, nbands - 1):
            l1 = lporder(f1[i - 1], f2[i - 1], devs[i], devs[i - 1])
            l2 = lporder(f1[i], f2[i], devs[i], devs[i + 1])
            l = max(l, l1, l2)
    n = int(math.ceil(l)) - 1
    ff = [0] + fcuts + [1]
    for i in range(1, len(ff) - 1):
        ff[i] *= 2
    aa = []
    for a in mags:
        aa = aa + [a, a]
    max_dev = max(devs)
    wts = [max_dev / d for d in devs]
    return n, ff, aa, wts


def low_pass(gain, Fs, freq1, freq2, passband_ripple_db, stopband_atten_db, nextra_taps=2):
    """optfir.low_pass (optfir.py:45-62)."""
    passband_dev = passband_ripple_to_dev(passband_ripple_db)
    stopband_dev = stopband_atten_to_dev(stopband_atten_db)
    n, fo, ao, w = remezord([freq1, freq2], (gain, 0), [passband_dev, stopband_dev], Fs)
    return remez(n + nextra_taps, fo, ao, w, "bandpass")


def high_pass(gain, Fs, freq1, freq2, passband_ripple_db, stopband_atten_db, nextra_taps=2):
    """optfir.high_pass (optfir.py:142-158): odd number of taps."""
    passband_dev = passband_ripple_to_dev(passband_ripple_db)
    stopband_dev = stopband_atten_to_dev(stopband_atten_db)
    n, fo, ao, w = remezord([freq1, freq2], (0, 1), [stopband_dev, passband_dev], Fs)
    if (n + nextra_taps) % 2 == 1:
        n += 1
    return remez(n + nextra_taps, fo, ao, w, "bandpass")


def band_pass(gain, Fs, freq_sb1, freq_pb1, freq_pb2, freq_sb2, passband_ripple_db, stopband_atten_db, nextra_taps=2):
    """optfir.band_pass (optfir.py:75-88)."""
    passband_dev = passband_ripple_to_dev(passband_ripple_db)
    stopband_dev = stopband_atten_to_dev(stopband_atten_db)
    n, fo, ao, w = remezord([freq_sb1, freq_pb1, freq_pb2, freq_sb2], (0, gain, 0),
                            [stopband_dev, passband_dev, stopband_dev], Fs)
    return remez(n + nextra_taps, fo, ao, w, "bandpass")


def pfb_channelizer_default_taps(numchans, atten=100):
    """The taps blks2.pfb_channelizer_ccf designs when none are given (blks2impl/pfb_channelizer.py:40-59): a low-pass
    over the full input band with bw 0.4, transition 0.2 (in units of the channel spacing), 0.1 dB of ripple, raised
    by 0.01 dB while the exchange does not converge."""
    bw, tb, ripple = 0.4, 0.2, 0.1
    while True:
        try:
            return low_pass(1, numchans, bw, bw + tb, ripple, atten)
        except RuntimeError:
            ripple += 0.01
            if ripple >= 1.0:
                raise RuntimeError("optfir could not generate an appropriate filter.")
