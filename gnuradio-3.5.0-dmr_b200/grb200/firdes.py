"""Host-side tap / window design, the gr_firdes subset the hot-path constructors are fed from
(gnuradio-core/src/lib/general/gr_firdes.cc:57-147,601-655,720-782).
float64 arithmetic evaluated exactly as the reference does, stored to float32."""
import math

import numpy as np

WIN_HAMMING, WIN_HANN, WIN_BLACKMAN, WIN_RECTANGULAR, WIN_KAISER, WIN_BLACKMAN_hARRIS = range(6)
WIN_BLACKMAN_HARRIS = WIN_BLACKMAN_hARRIS


def _izero(x):  # gr_firdes.cc:35-51
    s = u = 1.0
    n = 1
    halfx = x / 2.0
    while True:
        temp = halfx / n
        n += 1
        temp *= temp
        u *= temp
        s += u
        if not (u >= 1e-21 * s):
            return s


def window(win_type, ntaps, beta=6.76):
    """gr_firdes::window (:720-782) including its quirks: WIN_RECTANGULAR falls through into the
    Hamming case; the Blackman-harris loop leaves the last tap of an odd-length window at 0."""
    M = ntaps - 1
    n = np.arange(ntaps, dtype=np.float64)
    if win_type in (WIN_HAMMING, WIN_RECTANGULAR):
        w = 0.54 - 0.46 * np.cos((2 * math.pi * n) / M)
    elif win_type == WIN_HANN:
        w = 0.5 - 0.5 * np.cos((2 * math.pi * n) / M)
    elif win_type == WIN_BLACKMAN:
        w = 0.42 - 0.50 * np.cos((2 * math.pi * n) / (M - 1)) - 0.08 * np.cos((4 * math.pi * n) / (M - 1))
    elif win_type == WIN_BLACKMAN_hARRIS:
        w = np.zeros(ntaps)
        k = np.arange(-(ntaps // 2), ntaps // 2, dtype=np.float64)
        Mf = float(np.float32(M))
        w[: len(k)] = (0.35875 + 0.48829 * np.cos((2 * math.pi * k) / Mf) + 0.14128 * np.cos((4 * math.pi * k) / Mf) +
                       0.01168 * np.cos((6 * math.pi * k) / Mf))
    elif win_type == WIN_KAISER:
        ibeta = 1.0 / _izero(beta)
        inm1 = 1.0 / ntaps
        w = np.array([_izero(beta * math.sqrt(1.0 - (i * inm1) ** 2)) * ibeta for i in range(ntaps)])
    else:
        raise IndexError("gr_firdes:window: type out of range")
    return w.astype(np.float32)


def _sanity_1f(fs, fa, tw):  # :784-797
    if fs <= 0.0:
        raise IndexError("gr_firdes check failed: sampling_freq > 0")
    if fa <= 0.0 or fa > fs / 2:
        raise IndexError("gr_firdes check failed: 0 < fa <= sampling_freq / 2")
    if tw <= 0:
        raise IndexError("gr_firdes check failed: transition_width > 0")


def _low_pass_common(gain, fs, fc, ntaps, win_type, beta):
    w = window(win_type, ntaps, beta)
    M = (ntaps - 1) // 2
    fwT0 = 2 * math.pi * fc / fs
    taps = np.zeros(ntaps, np.float32)
    for n in range(-M, M + 1):
        if n == 0:
            taps[n + M] = fwT0 / math.pi * float(w[n + M])
        else:
            taps[n + M] = math.sin(n * fwT0) / (n * math.pi) * float(w[n + M])
    fmax = float(taps[M])
    for n in range(1, M + 1):
        fmax += 2 * float(taps[n + M])
    g = gain / fmax
    return np.array([np.float32(float(t) * g) for t in taps], np.float32)


def low_pass(gain, sampling_freq, cutoff_freq, transition_width, window_type=WIN_HAMMING, beta=6.76):
    """gr_firdes::low_pass (:104-147)."""
    _sanity_1f(sampling_freq, cutoff_freq, transition_width)
    width_factor = [np.float32(3.3), np.float32(3.1), np.float32(5.5), np.float32(2.0), np.float32(10.0)]
    if not 0 <= window_type < len(width_factor):
        raise IndexError("gr_firdes::low_pass: window type has no width factor (use low_pass_2)")
    ntaps = int(float(width_factor[window_type]) / (transition_width / sampling_freq) + 0.5)
    if ntaps & 1 == 0:
        ntaps += 1
    return _low_pass_common(gain, sampling_freq, cutoff_freq, ntaps, window_type, beta)


def low_pass_2(gain, sampling_freq, cutoff_freq, transition_width, attenuation_dB, window_type=WIN_HAMMING, beta=6.76):
    """gr_firdes::low_pass_2 (:57-101)."""
    _sanity_1f(sampling_freq, cutoff_freq, transition_width)
    ntaps = int(attenuation_dB * sampling_freq / (22.0 * transition_width))
    if ntaps & 1 == 0:
        ntaps += 1
    return _low_pass_common(gain, sampling_freq, cutoff_freq, ntaps, window_type, beta)


def root_raised_cosine(gain, sampling_freq, symbol_rate, alpha, ntaps):
    """gr_firdes::root_raised_cosine (:601-655)."""
    ntaps |= 1
    spb = sampling_freq / symbol_rate
    taps = np.zeros(ntaps, np.float32)
    scale = 0.0
    for i in range(ntaps):
        xindx = float(i - ntaps // 2)
        x1 = math.pi * xindx / spb
        x2 = 4 * alpha * xindx / spb
        x3 = x2 * x2 - 1
        if abs(x3) >= 0.000001:
            if i != ntaps // 2:
                num = math.cos((1 + alpha) * x1) + math.sin((1 - alpha) * x1) / (4 * alpha * xindx / spb)
            else:
                num = math.cos((1 + alpha) * x1) + (1 - alpha) * math.pi / (4 * alpha)
            den = x3 * math.pi
        else:
            if alpha == 1:
                taps[i] = -1
                continue
            x3 = (1 - alpha) * x1
            x2 = (1 + alpha) * x1
            num = (math.sin(x2) * (1 + alpha) * math.pi - math.cos(x3) * ((1 - alpha) * math.pi * spb) / (4 * alpha * xindx) +
                   math.sin(x3) * spb * spb / (4 * alpha * xindx * xindx))
            den = -32 * math.pi * alpha * alpha * xindx / spb
        taps[i] = 4 * alpha * num / den
        scale += float(taps[i])
    return np.array([np.float32(float(t) * gain / scale) for t in taps], np.float32)
