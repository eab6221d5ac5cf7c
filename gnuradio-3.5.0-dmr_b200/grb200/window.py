"""gnuradio.window (gnuradio-core/src/python/gnuradio/window.py:152-180): the cosine-sum windows, evaluated at
(index + 0.5) / (fft_size - 1) like the reference (NOT the textbook index / (N - 1), and not gr_firdes::window either:
gr_firdes.cc:744-748 -- the two Blackman-Harris windows of the tree differ slightly)."""
import math


def coswindow(coeffs):
    def closure(fft_size):
        window = [0.0] * fft_size
        for w_index in range(fft_size):
            for c_index, coeff in enumerate(coeffs):
                window[w_index] += (-1) ** c_index * coeff * math.cos(2.0 * c_index * math.pi * (w_index + 0.5) / (fft_size - 1))
        return window
    return closure


blackmanharris = coswindow((0.35875, 0.48829, 0.14128, 0.01168))
nuttall = coswindow((0.3635819, 0.4891775, 0.1365995, 0.0106411))
nuttall_cfd = coswindow((0.355768, 0.487396, 0.144232, 0.012604))
flattop = coswindow((1.0, 1.93, 1.29, 0.388, 0.032))


def rectangular(fft_size):
    return [1] * fft_size
