#!/usr/bin/env python3
"""Stand-alone timing of the clock-recovery kernel (mm_ws_kernel) in each of its builds, on the matched-filter output
of the cfg5 workload (8000 channels, 800 of them carrying DMR bursts, the rest noise).

The front of the chain (channelizer -> discriminator + matched filter) is run ONCE through the device entry points of
the blocks to produce F [rows][M]; then every kernel variant (grcuda_clock_recovery_mm_ff_set_kernel_variant) processes
the same F from the same initial state, alone on the device, timed with CUDA events; soft symbols, slicer decisions,
symbol counts and the final loop state of every variant must be bit identical to variant 0's.

  python tools/mm_microbench.py [--rows 12500] [--reps 5] [--variants 0,1,2,...] [--json gpurun_out/mm_microbench.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=12500)
    ap.add_argument("--active", type=int, default=800)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--variants", default="0,1,2,3,4,5,6,7,8,9")
    ap.add_argument("--json", default="")
    ap.add_argument("--chain", action="store_true", help="time the FUSED tail (clock recovery + slicer + map + correlator) of the "
                    "flagship chain instead of the stand-alone block: front and tail of each block on one stream, events around the tail")
    ap.add_argument("--stats", action="store_true", help="lab build (make EXTRA=-DMMW_STATS, GRCUDA_LIB=...): core-warp cycle counters")
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from grb200 import blocks, lib, synth_torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    M, R = bench.M, args.rows
    cfg = bench.chain_config(R)
    if args.chain:
        return chain_mode(args, torch, np, bench, lib, synth_torch, dev)
    pfb = blocks.pfb_channelizer_ccf(M, cfg.pfb_taps)
    quad = blocks.quadrature_demod_cf(cfg.quad_gain)
    rrc = blocks.fir_filter_fff(1, cfg.rrc_taps)
    Th = pfb.history() - 1
    YH = blocks.quad_demod_fir_fff_history(rrc)
    x, _ = synth_torch.wideband_block(M, R, Th, args.active, 1234, dev)
    Y = torch.zeros((YH + R, M), dtype=torch.complex64, device=dev)
    pfb.work_device(R, x, Y[YH:])
    del x
    KEEP = 64
    F = torch.zeros((KEEP + R, M), dtype=torch.float32, device=dev)
    blocks.quad_demod_fir_fff_work_device(quad, rrc, R, M, Y, F[KEEP:], 0)
    torch.cuda.synchronize()
    del Y
    ninput = KEEP + R
    max_out = int(np.ceil(ninput / cfg.omega * 1.25)) + 64
    sym_per_chan = None
    ref = None
    results = []
    clk = None
    for v in [int(t) for t in args.variants.split(",")]:
        out = torch.zeros((max_out, M), dtype=torch.float32, device=dev)
        sl = torch.zeros((max_out, M), dtype=torch.uint8, device=dev)
        cnt = torch.zeros((M,), dtype=torch.int32, device=dev)
        times = []
        st = None
        for rep in range(args.reps + 1):
            mm = blocks.clock_recovery_mm_ff(cfg.omega, cfg.gain_omega, cfg.mu, cfg.gain_mu, cfg.omega_relative_limit, nchan=M)
            mm.set_slicer(4, 0.0)
            mm.set_kernel_variant(v)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            mm.work_device(ninput, -KEEP, F, out, sl, max_out, cnt)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                times.append(e0.elapsed_time(e1))
            if rep == args.reps:
                st = [mm._state(c) for c in range(0, M, 97)]
            if args.stats:
                import ctypes as C
                buf = (C.c_ulonglong * 12)()
                lib.load().grcuda_lab_mm_stats(buf)
                stats = list(buf)
            del mm
        c = cnt.cpu().numpy()
        o = out.cpu().numpy()
        s = sl.cpu().numpy()
        mask = np.arange(max_out)[:, None] < c[None, :]
        got = (c.copy(), np.where(mask, o.view(np.uint32), 0), np.where(mask, s, 0), st)
        same = None
        if ref is None:
            ref = got
            sym_per_chan = float(c.mean())
        else:
            same = bool(np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]) and np.array_equal(ref[2], got[2])
                        and ref[3] == got[3])
        best = min(times)
        # cycles per symbol of the slowest warp ~ kernel time x SM clock / symbols per channel
        r = {"variant": v, "ms_best": best, "ms_all": times, "identical_to_first": same,
             "symbols_per_channel": sym_per_chan, "ns_per_symbol": best * 1e6 / sym_per_chan,
             "cycles_per_symbol_at_1965MHz": best * 1e-3 * 1.965e9 / sym_per_chan}
        if args.stats and stats[7]:
            w = stats[7]
            r["core_stats_last_rep"] = {
                "full_trips_per_warp": stats[1] / w, "cycles_per_full_trip": stats[0] / max(stats[1], 1),
                "other_trips_per_warp": stats[3] / w, "cycles_per_other_trip": stats[2] / max(stats[3], 1),
                "single_sections_per_warp": stats[5] / w, "cycles_per_single_section": stats[4] / max(stats[5], 1),
                "core_cycles_per_warp": stats[6] / w, "lane_trips_blocked_by_queue_per_warp": stats[8] / w,
                "lane_trips_blocked_by_input_per_warp": stats[9] / w, "lane_trips_stopped_half_way_per_warp": stats[10] / w,
                "raw": stats}
        results.append(r)
        print(json.dumps(r), flush=True)
        del out, sl, cnt
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        json.dump(results, open(args.json, "w"), indent=1)
    bad = [r["variant"] for r in results if r["identical_to_first"] is False]
    if bad:
        print("MISMATCH in variants", bad)
        return 1
    return 0


def chain_mode(args, torch, np, bench, lib, synth_torch, dev):
    """Every variant runs the same 3 blocks through its own chain; the tail kernel is timed alone on the device; the
    sync-hit list, symbol counts and soft symbols of the last block must equal variant 0's."""
    from grb200 import chain
    M, R = bench.M, args.rows
    ref = None
    results = []
    for v in [int(t) for t in args.variants.split(",")]:
        ch = chain.DmrChain(bench.chain_config(R))
        ch.set_tail_variant(v % 100)
        ch.set_split_correlator(v >= 100)    # variant + 100: correlator as its own kernel
        Th = ch.history_rows()
        x, _ = synth_torch.wideband_block(M, R, Th, args.active, 1234, dev)
        stream = torch.cuda.current_stream().cuda_stream
        times = []
        stats = None
        for rep in range(args.reps + 1):
            ch.process_front_device(x, R, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ch.process_tail_device(stream)
            e1.record()
            torch.cuda.synchronize()
            if rep:
                times.append(e0.elapsed_time(e1))
            if args.stats:
                import ctypes as C
                buf = (C.c_ulonglong * 12)()
                lib.load().grcuda_lab_mm_stats(buf)
                stats = list(buf)
        res = ch.fetch()
        hits, nh = ch.read_hits_array(1 << 20)
        hl = sorted((int(h["channel"]), int(h["bit_index"])) for h in hits[:nh]) if nh else []
        c = res["counts"]
        mask = np.arange(res["soft"].shape[0])[:, None] < c[None, :]
        got = (c.copy(), np.where(mask, res["soft"].view(np.uint32), 0), np.where(mask, res["symbols"], 0), hl)
        same = None
        if ref is None:
            ref = got
        else:
            same = bool(np.array_equal(ref[0], got[0]) and np.array_equal(ref[1], got[1]) and np.array_equal(ref[2], got[2])
                        and ref[3] == got[3])
        best = min(times)
        spc = float(c.mean())
        r = {"variant": v, "mode": "chain tail (fused)", "ms_best": best, "ms_all": times, "identical_to_first": same, "sync_hits": nh,
             "symbols_per_channel": spc, "cycles_per_symbol_at_1965MHz": best * 1e-3 * 1.965e9 / spc}
        if stats and stats[7]:
            w = stats[7]
            r["core_stats_last_rep"] = {
                "full_trips_per_warp": stats[1] / w, "cycles_per_full_trip": stats[0] / max(stats[1], 1),
                "other_trips_per_warp": stats[3] / w, "cycles_per_other_trip": stats[2] / max(stats[3], 1),
                "single_sections_per_warp": stats[5] / w, "cycles_per_single_section": stats[4] / max(stats[5], 1),
                "core_cycles_per_warp": stats[6] / w, "lane_trips_blocked_by_queue_per_warp": stats[8] / w,
                "lane_trips_blocked_by_input_per_warp": stats[9] / w, "lane_trips_stopped_half_way_per_warp": stats[10] / w,
                "raw": stats}
        results.append(r)
        print(json.dumps(r), flush=True)
        del ch, x
    if args.json:
        os.makedirs(os.path.dirname(args.json) or ".", exist_ok=True)
        json.dump(results, open(args.json, "w"), indent=1)
    bad = [r["variant"] for r in results if r["identical_to_first"] is False]
    if bad:
        print("MISMATCH in variants", bad)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
