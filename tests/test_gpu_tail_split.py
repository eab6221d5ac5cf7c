"""The tail of the flagship chain as two kernels (clock recovery + slicer, then the time-parallel access-code
correlator) against the fused kernel, and a chain driven the way a time shard drives it (seek to -halo, front on
halo + R rows, external state buffers chained from block to block) against one continuous stream: sync hits, symbol
counts, soft symbols and the exported loop state must be identical."""
import numpy as np
import pytest

from conftest import has_cuda

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not has_cuda(), reason="needs a CUDA device")]

M, T = 160, 16


def make_chain(rows):
    from grb200 import chain, firdes
    taps = firdes.low_pass_2(1.0, M * 12500.0, 5500.0, 1500.0, 60.0, firdes.WIN_BLACKMAN_hARRIS)
    c = len(taps) // 2
    taps = (taps[c - M * T // 2: c - M * T // 2 + M * T] * M).astype(np.float32)
    return chain.DmrChain(chain.DmrChainConfig(M, taps, max_rows_per_block=rows))


def stream(nrows, seed=5):
    from grb200 import synth
    rng = np.random.default_rng(seed)
    active = sorted(set(int(c) for c in rng.choice(M, size=20, replace=False)))
    x, _ = synth.wideband_compose(rng, M, nrows, active, noise_sigma=3e-3)
    return x.reshape(nrows, M), active


def test_split_tail_equals_fused_tail():
    import torch
    R, nb = 1400, 4
    xr, _ = stream(R * nb)
    res = {}
    for split in (False, True):
        ch = make_chain(R)
        ch.set_split_correlator(split)
        Th = ch.history_rows()
        buf = torch.from_numpy(np.concatenate([np.zeros((Th, M), np.complex64), xr])).cuda()
        st = torch.zeros(ch.state_bytes(), dtype=torch.uint8, device="cuda")
        out = []
        s0 = torch.cuda.current_stream().cuda_stream
        for b in range(nb):
            ch.process_front_device(buf[b * R:], R, s0)
            ch.process_tail_device(s0)
            r = ch.fetch()
            hits, nh = ch.read_hits()
            out.append((r["counts"].copy(), r["symbols"].copy(), sorted(hits)))
        ch.export_state(st, s0)
        torch.cuda.synchronize()
        assert ch.counters() == {"clamped": 0, "overflow": 0, "hits_dropped": 0}
        res[split] = (out, st.cpu().numpy())
    tot = 0
    for b in range(nb):
        f, s = res[False][0][b], res[True][0][b]
        assert np.array_equal(f[0], s[0]), b
        m = np.arange(f[1].shape[0])[:, None] < f[0][None, :]
        assert np.array_equal(np.where(m, f[1], 0), np.where(m, s[1], 0)), b
        assert f[2] == s[2], b
        tot += len(f[2])
    assert tot > 50
    assert np.array_equal(res[False][1], res[True][1])      # loop state incl. the correlator registers and bit counts


def test_shard_style_driving_equals_continuous_stream():
    import torch
    R, nb = 1500, 3
    xr, active = stream(R * nb, seed=6)
    probe = make_chain(512)
    halo, Th = probe.warmup_rows(), probe.history_rows()
    del probe
    H = Th + halo
    full = torch.from_numpy(np.concatenate([np.zeros((H, M), np.complex64), xr])).cuda()   # row H = stream row 0
    s0 = torch.cuda.current_stream().cuda_stream
    a = make_chain(R + halo)
    u8 = dict(dtype=torch.uint8, device="cuda")
    mm_buf = [torch.zeros(a.mm_state_bytes(), **u8) for _ in range(2)]
    co_buf = [torch.zeros(a.corr_state_bytes(), **u8) for _ in range(2)]
    got = []
    for b in range(nb):
        a.seek_async(b * R - halo, s0)
        a.process_front_device(full[b * R:], halo + R, s0)            # Th history rows + halo + R
        a.process_tail_mm_device(mm_buf[(b + 1) % 2] if b else None, mm_buf[b % 2], s0)
        a.process_tail_corr_device(co_buf[(b + 1) % 2] if b else None, co_buf[b % 2], s0)
        r = a.fetch()
        hits, _ = a.read_hits()
        got.append((r["counts"].copy(), r["soft"].copy(), sorted(hits)))
    assert a.counters() == {"clamped": 0, "overflow": 0, "hits_dropped": 0}
    c = make_chain(R + halo)
    tot = 0
    for b in range(nb):
        if b == 0:
            c.seek_async(-halo, s0)
            c.process_front_device(full, halo + R, s0)
        else:
            c.process_front_device(full[halo + b * R:], R, s0)
        c.process_tail_device(s0)
        r = c.fetch()
        hits, _ = c.read_hits()
        assert np.array_equal(r["counts"], got[b][0]), b
        m = np.arange(r["soft"].shape[0])[:, None] < r["counts"][None, :]
        assert np.array_equal(np.where(m, r["soft"].view(np.uint32), 0), np.where(m, got[b][1].view(np.uint32), 0)), b
        assert sorted(hits) == got[b][2], b
        tot += len(hits)
    assert tot >= len(active)
