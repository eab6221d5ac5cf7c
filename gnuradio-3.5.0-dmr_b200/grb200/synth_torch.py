"""Synthetic wideband input generated ON the device with torch (bench.py / large GPU tests).

Channel-rate 4FSK basebands for the active channels are built with torch ops, then a crude
polyphase SYNTHESIS (one inverse FFT per row across channels, rectangular prototype) turns the
[rows][M] channel matrix into the interleaved wideband stream x[m*M + j].  torch.fft here only
fabricates test input; it is not on the measured path."""
import math

import numpy as np

from . import synth


def dmr_frequency_tracks(rows, n_active, fs_channel, seed, device):
    """[n_active][rows] instantaneous-frequency tracks (units of the symbol levels +-1, +-3), 144-symbol
    slots with the 24-symbol BS-data sync in the middle, RRC-shaped at fs_channel/4800 samples per symbol."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    sps = fs_channel / synth.SYMBOL_RATE
    nsym = int(rows / sps) + 16
    nslots = nsym // 144 + 1
    sym = (torch.randint(0, 4, (n_active, nslots * 144), generator=g) * 2 - 3).to(torch.float32)
    sync = torch.tensor(synth.bits_to_symbols(synth.DMR_BS_DATA_SYNC_BITS), dtype=torch.float32)
    off = torch.randint(0, 144, (n_active,), generator=g)
    for s in range(nslots):
        sym[:, s * 144 + 54: s * 144 + 78] = sync
    sym = sym.to(device)
    # shaped[n] = sum_k a_k h(n/sps - k): evaluate the continuous RRC on a +-6 symbol window
    n = torch.arange(rows, device=device, dtype=torch.float64)
    t = n / sps
    k0 = torch.floor(t).to(torch.int64)
    out = torch.zeros((n_active, rows), device=device, dtype=torch.float32)
    a = synth.RRC_ALPHA
    for j in range(-6, 7):
        k = (k0 + j).clamp(0, sym.shape[1] - 1)
        tau = (t - (k0 + j).to(torch.float64))
        tau = torch.where(tau.abs() < 1e-9, torch.full_like(tau, 1e-9), tau)
        tau = torch.where((tau.abs() - 1.0 / (4 * a)).abs() < 1e-9, tau + 1e-7, tau)
        h = (torch.sin(math.pi * tau * (1 - a)) + 4 * a * tau * torch.cos(math.pi * tau * (1 + a))) / \
            (math.pi * tau * (1 - (4 * a * tau) ** 2))
        out += sym[:, k] * h.to(torch.float32)[None, :]
    # per-channel circular time offset so that bursts are not aligned across channels
    idx = (torch.arange(rows, device=device)[None, :] + (off.to(device) * 2)[:, None]) % rows
    return torch.gather(out, 1, idx)


def wideband_block(M, rows, history_rows, n_active, seed, device, noise_sigma=1e-3, fs_channel=synth.CHANNEL_SPACING):
    """Returns (x [history_rows + rows][M] complex64 on `device`, active channel indices).
    The first history_rows rows are a circular continuation (the block tiles seamlessly in time)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    active = torch.randperm(M, generator=g)[:n_active].sort().values
    freq = dmr_frequency_tracks(rows, n_active, fs_channel, seed, device)
    phase = torch.cumsum(freq.to(torch.float64) * (2 * math.pi * synth.DEVIATION_HZ / fs_channel), dim=1)
    base = torch.polar(torch.ones_like(phase, dtype=torch.float32), phase.to(torch.float32))  # [n_active][rows]
    chan = torch.zeros((rows, M), dtype=torch.complex64, device=device)
    chan[:, active.to(device)] = base.transpose(0, 1)
    # synthesis: x[m][j] = sum_c chan[m][c] e^{+2 pi i c j / M}  (channel c centred at c*fs/M)
    x = torch.fft.ifft(chan, dim=1, norm="forward")
    del chan
    gen = torch.Generator(device=device).manual_seed(seed + 2)
    x += noise_sigma * torch.view_as_complex(torch.randn((rows, M, 2), generator=gen, device=device, dtype=torch.float32))
    out = torch.empty((history_rows + rows, M), dtype=torch.complex64, device=device)
    out[history_rows:] = x
    out[:history_rows] = x[rows - history_rows:]
    return out, active.numpy()
