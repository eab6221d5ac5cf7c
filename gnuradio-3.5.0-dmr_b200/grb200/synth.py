"""Synthetic DMR-like test signals (numpy; no reference or oracle dependency).

DMR air interface facts used here (ETSI TS 102 361-1): 4FSK at 4800 symbol/s in 12.5 kHz
channels, RRC alpha = 0.2, deviations +-648 Hz / +-1944 Hz for symbols +-1 / +-3, dibit map
01 -> +3, 00 -> +1, 10 -> -1, 11 -> -3, 48-bit sync patterns in the middle of each 27.5 ms burst,
one burst per 30 ms TDMA slot (144 symbols).  The reference tree contains no DMR code (SURVEY.md
section 0), so these constants are inputs to the generic blocks, not part of the parity contract.
"""
import numpy as np

SYMBOL_RATE = 4800.0
CHANNEL_SPACING = 12500.0
DEVIATION_HZ = 648.0  # per unit symbol level
RRC_ALPHA = 0.2

DMR_SYNC_WORDS = {
    "bs_voice": 0x755FD7DF75F7,
    "bs_data": 0xDFF57D75DF5D,
    "ms_voice": 0x7F7D5DD57DFD,
    "ms_data": 0xD5D7F77FD757,
}


def word_bits(word, nbits=48):
    return [(word >> (nbits - 1 - i)) & 1 for i in range(nbits)]


DMR_BS_DATA_SYNC_BITS = word_bits(DMR_SYNC_WORDS["bs_data"])
DMR_BS_VOICE_SYNC_BITS = word_bits(DMR_SYNC_WORDS["bs_voice"])

# dibit (b1 b0 as integer 0..3) -> symbol level
DIBIT_TO_SYMBOL = {0b01: 3, 0b00: 1, 0b10: -1, 0b11: -3}
# 4-level slicer decision (0..3 for levels -3,-1,+1,+3; pager_slicer_fb.cc:47-69) -> dibit value
SLICER_TO_DIBIT_MAP = [0b11, 0b10, 0b00, 0b01]


def access_code_string(bits):
    """'1'/'0' string as taken by digital_make_correlate_access_code_bb."""
    return "".join("1" if b else "0" for b in bits)


def bits_to_symbols(bits):
    bits = np.asarray(bits, dtype=np.int64)
    d = bits[0::2] * 2 + bits[1::2]
    lut = np.zeros(4, np.int64)
    for k, v in DIBIT_TO_SYMBOL.items():
        lut[k] = v
    return lut[d]


def dmr_symbols(rng, nslots, sync_bits=None, idle_fraction=0.0):
    """nslots TDMA slots of 144 symbols: 54 payload, 24 sync, 54 payload, 12 guard symbols.
    Returns (symbols, sync_start_symbol_indices)."""
    sync_bits = DMR_BS_DATA_SYNC_BITS if sync_bits is None else sync_bits
    sync = bits_to_symbols(sync_bits)
    out = np.zeros(nslots * 144, np.int64)
    starts = []
    for s in range(nslots):
        base = s * 144
        if idle_fraction and rng.random() < idle_fraction:
            out[base:base + 144] = rng.integers(0, 4, 144) * 2 - 3
            continue
        out[base:base + 54] = rng.integers(0, 4, 54) * 2 - 3
        out[base + 54:base + 78] = sync
        out[base + 78:base + 132] = rng.integers(0, 4, 54) * 2 - 3
        out[base + 132:base + 144] = rng.integers(0, 4, 12) * 2 - 3
        starts.append(base + 54)
    return out, np.array(starts, np.int64)


def rrc_pulse(t, alpha=RRC_ALPHA):
    """Continuous root-raised-cosine impulse response, T = 1 symbol, unit energy scaling."""
    t = np.asarray(t, np.float64)
    out = np.empty_like(t)
    eps = 1e-9
    z = np.abs(t) < eps
    q = np.abs(np.abs(t) - 1.0 / (4 * alpha)) < eps
    r = ~(z | q)
    tr = t[r]
    out[r] = (np.sin(np.pi * tr * (1 - alpha)) + 4 * alpha * tr * np.cos(np.pi * tr * (1 + alpha))) / \
             (np.pi * tr * (1 - (4 * alpha * tr) ** 2))
    out[z] = 1 - alpha + 4 * alpha / np.pi
    out[q] = alpha / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * alpha)) +
                                   (1 - 2 / np.pi) * np.cos(np.pi / (4 * alpha)))
    return out


def shape_symbols(symbols, sps, rx_taps=None, span=8, alpha=RRC_ALPHA, nsamples=None):
    """Transmit pulse shaping at a (possibly non-integer) number of samples per symbol:
    s[n] = sum_k a_k h(n/sps - k) with the continuous RRC pulse.  rx_taps is unused (kept so
    callers can pass the matched filter for documentation)."""
    symbols = np.asarray(symbols, np.float64)
    n = int(len(symbols) * sps) + 1 if nsamples is None else nsamples
    t = np.arange(n) / sps
    k0 = np.floor(t).astype(np.int64)
    s = np.zeros(n)
    for j in range(-span, span + 1):
        k = k0 + j
        ok = (k >= 0) & (k < len(symbols))
        s[ok] += symbols[k[ok]] * rrc_pulse(t[ok] - k[ok], alpha)
    return s


def fm_modulate(freq_units, fs, deviation_hz=DEVIATION_HZ, phase0=0.0):
    """Complex baseband FM: instantaneous frequency = deviation_hz * freq_units[n]."""
    ph = phase0 + 2 * np.pi * deviation_hz / fs * np.cumsum(np.asarray(freq_units, np.float64))
    return np.exp(1j * ph)


def dmr_channel_baseband(rng, nslots, fs, snr_db=None, sync_bits=None, amplitude=1.0):
    """One 4FSK channel at sample rate fs: returns (complex64 samples, symbols, sync_starts)."""
    sym, starts = dmr_symbols(rng, nslots, sync_bits)
    shaped = shape_symbols(sym, fs / SYMBOL_RATE)
    x = amplitude * fm_modulate(shaped, fs)
    if snr_db is not None:
        sigma = amplitude * 10 ** (-snr_db / 20) / np.sqrt(2)
        x = x + sigma * (rng.standard_normal(len(x)) + 1j * rng.standard_normal(len(x)))
    return x.astype(np.complex64), sym, starts


def wideband_compose(rng, M, rows, active, fs_channel=CHANNEL_SPACING, noise_sigma=1e-3, amplitude=1.0,
                     sync_bits=None):
    """Wideband stream of rows*M samples at fs = M*fs_channel holding one DMR-like FM channel at
    each channel index in `active` (centre frequency c*fs/M, the PFB's bin c) plus white noise.
    The FM phase is computed at the channel rate and linearly interpolated to the wideband rate
    (the modulating signal is band-limited to ~3 kHz, so this is accurate to well below the
    noise floor used here).  Returns (x complex64, dict channel -> (symbols, sync_starts))."""
    n = rows * M
    t = np.arange(n, dtype=np.float64)
    x = noise_sigma * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    truth = {}
    nslots = int(np.ceil(rows / fs_channel * SYMBOL_RATE / 144)) + 1
    for c in active:
        sym, starts = dmr_symbols(rng, nslots, sync_bits)
        shaped = shape_symbols(sym, fs_channel / SYMBOL_RATE, nsamples=rows + 2)
        ph = 2 * np.pi * DEVIATION_HZ / fs_channel * np.cumsum(shaped)
        ph_w = np.interp(t / M, np.arange(rows + 2), ph)
        x += amplitude * np.exp(1j * (ph_w + 2 * np.pi * ((c * t) % M) / M))
        truth[c] = (sym, starts)
    return x.astype(np.complex64), truth
