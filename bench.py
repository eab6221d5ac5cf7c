#!/usr/bin/env python3
"""bench.py -- input MS/s of the wideband PFB channelizer + batched DMR demod on 1/2/4/8 B200.

Workload (BASELINE.json configs[4], named in `config.workload`): 100 MS/s complex stream -> 8000 x
12.5 kHz channels (gr_pfb_channelizer_ccf, 16 taps/branch) -> per channel quadrature_demod_cf ->
RRC fir_filter_fff -> clock_recovery_mm_ff -> 4-level slicer -> map/unpack -> correlate_access_code_bb.
One "step" = one pass of the whole path over one time block of `--rows` channel-rate rows
(default 12500 rows = 1 s of signal = 100 M input samples = 800 MB, i.e. larger than L2) per GPU.
N > 1: one process per GPU (torchrun), the stream is time sharded (block = step*N + rank), halos and
the per-channel loop state travel by NCCL send/recv (grb200/sharding.py).  Weak scaling.

Prints ONE JSON line (rank 0).  `--impl reference` times the reference's own CPU implementation of
the same path (oracle/_ref: the reference sources compiled in place) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))

M = 8000
T = 16
FS_CHANNEL = 12500.0
WORKLOAD = "cfg5: 100 MS/s -> 8000 x 12.5 kHz PFB channelizer (16 taps/branch) + batched DMR 4FSK demod + sync search"
SYMBOLS_PER_ROW = 4800.0 / FS_CHANNEL   # 4800 baud at 12.5 kS/s: one symbol per 2.604 rows
# Per stage: algorithmic HBM bytes per input sample (SURVEY 8d / DESIGN.md section 4) and the roofline that really bounds
# it.  "rrc_fir" is the chain's name for the FUSED discriminator + matched filter kernel (demod_front_kernel:
# 8 B of channelizer output in, 4 B of filtered soft samples out = 12 B); the stand-alone discriminator stage
# ("quad_demod") only runs in the unfused lab build.
STAGE_BYTES = {
    "pfb_fir": 16.0, "pfb_fft": 16.0, "quad_demod": 12.0, "rrc_fir": 12.0,
    "mm_slicer": 4.0 + (4.0 + 1.0) * SYMBOLS_PER_ROW, "map_unpack_corr": (1.0 + 2.0) * SYMBOLS_PER_ROW,
}
STAGE_KERNEL = {"pfb_fir": "pfb_fir_tma_kernel", "pfb_fft": "fft_fixed_kernel", "quad_demod": "quad_demod_kernel",
                "rrc_fir": "demod_front_kernel (quadrature_demod_cf + fir_filter_fff fused)",
                "mm_slicer": "mm_quad_kernel / mm_ws_kernel (clock_recovery_mm_ff + slicer)",
                "map_unpack_corr": "corr_par_kernel (map_bb + unpack_k_bits_bb + correlate_access_code_bb)"}
STAGE_BOUND = {"pfb_fir": "hbm", "pfb_fft": "hbm", "quad_demod": "hbm", "rrc_fir": "fp32_issue", "mm_slicer": "latency",
               "map_unpack_corr": "issue"}
# The clock-recovery recursion: the dependent chain from one symbol's mu to the next (DESIGN.md section 4 lists the
# operations): 23 dependent FP32/integer operations + one shared-memory round trip, at the latencies measured on this
# pool's B200 (profiles/r2_fp32_peaks.json: 4.11 cycles per dependent FADD/FMUL/FFMA, 23 per dependent LDS).
MM_CHAIN_OPS, MM_CHAIN_LDS = 23, 1


ROOFLINE_NOTES = {   # which roofline really bounds each stage (ncu evidence under profiles/)
    "pfb_fir": "HBM stream (TMA staged)",
    "pfb_fft": "HBM and latency at one 400-thread CTA per SM (46 instructions per point; ncu: issue active 32 %, dram 46 %)",
    "rrc_fir": "FP32-issue bound, not HBM bound: the reference's SSE summation order forbids FMA (separate IEEE multiply "
               "and add per tap) and the table arctangent needs a correctly rounded division; DRAM traffic = algorithmic bytes",
    "mm_slicer": "latency bound: one sequential recursion per channel (8000 channels = 250 warps at single-warp "
                 "instruction latency); its own roofline is the dependent-chain floor, see own_bound",
    "map_unpack_corr": "integer issue (popcount windows over packed dibits), off the serial chain",
}


def measured_fp32():
    """FP32 / issue / latency peaks measured on this pool's B200 by tools/measure_fp32_peaks.cu (committed result)."""
    p = os.path.join(ROOT, "profiles", "r2_fp32_peaks.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


def chain_config(max_rows, keep_bytes=False):
    import numpy as np
    from grb200 import chain, firdes
    fs = M * FS_CHANNEL
    # ~128 000-tap prototype (T = 16 taps per branch): low_pass_2, Blackman-harris, 60 dB
    taps = firdes.low_pass_2(float(M), fs, 5400.0, 2131.0, 60.0, firdes.WIN_BLACKMAN_hARRIS)
    assert (len(taps) + M - 1) // M == T, len(taps)
    return chain.DmrChainConfig(M, np.asarray(taps, np.float32), fs_channel=FS_CHANNEL, max_rows_per_block=max_rows,
                                keep_bytes=keep_bytes)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(",") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            r = [c.strip() for c in r]
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            sm.sort()
            out["sm_mhz"] = sm[len(sm) // 2]
        out["reasons"] = sorted(reasons)
        out["samples"] = len(sm)
        return out


def pin_to_gpu_numa_node(local):
    """Binds this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned allocation: first-touch then
    places the staging buffers there, and N ranks stop sharing one node's memory controllers and PCIe root."""
    try:
        online = open("/sys/devices/system/node/online").read().strip()
        if online in ("0", ""):
            return {"nodes_online": 1, "note": "single NUMA node: nothing to bind"}
        import torch
        uuid = str(torch.cuda.get_device_properties(local).uuid).lower()
        out = subprocess.run(["nvidia-smi", "--query-gpu=uuid,pci.bus_id", "--format=csv,noheader"], stdout=subprocess.PIPE,
                             stderr=subprocess.DEVNULL, text=True, timeout=20).stdout
        bus = None
        for line in out.strip().splitlines():
            u, _, b = line.partition(",")
            if u.strip().lower().endswith(uuid):
                bus = b.strip().lower()
        if bus is None:       # (UUIDs are redacted on some boxes) fall back to the enumeration order
            lines = [l for l in out.strip().splitlines() if l.strip()]
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[local]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else local
            if idx < len(lines):
                bus = lines[idx].partition(",")[2].strip().lower()
        if bus is None:
            return None
        if len(bus.split(":")[0]) == 8:       # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bus = bus[4:]
        path = "/sys/bus/pci/devices/%s/numa_node" % bus
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return {"node": node, "cpus": len(cpus)}
    except Exception:
        pass
    return None


def run_reference(args):
    """The reference arm: the reference's own blocks (oracle/_ref/libgrref.so) on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import refharness as R
    from grb200 import synth
    kind = "reference"
    if not R.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgrref.so was not built (needs /root/reference at build time)"}))
        return 0
    cfg = chain_config(args.cpu_rows)
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(6)
    rows = args.cpu_rows                                   # one step = one bounded sample of the workload
    x, _ = synth.wideband_compose(rng, M, min(rows, 64), [], noise_sigma=1.0)  # white noise rows (cost is data independent)
    x = np.tile(x, rows // min(rows, 64) + 1)[: rows * M]
    kw = dict(M=M, pfb_taps=cfg.pfb_taps, quad_gain=cfg.quad_gain, rrc_taps=cfg.rrc_taps, omega=cfg.omega,
              gain_omega=cfg.gain_omega, mu=cfg.mu, gain_mu=cfg.gain_mu, limit=cfg.omega_relative_limit,
              slicer_alpha=cfg.slicer_alpha, symbol_map=cfg.symbol_map, access_code=cfg.access_code, threshold=cfg.threshold,
              x=x, nthreads=cores, fft_fast=True)
    for _ in range(args.warmup_ref):
        R.bench_chain(**kw)
    tot = 0.0
    for _ in range(args.steps_ref):
        s0, s1, _, _ = R.bench_chain(**kw)
        tot += s0 + s1
    ms_step = tot / args.steps_ref * 1e3
    value = rows * M / (tot / args.steps_ref) / 1e6
    sample = ("%d rows x %d channels = %.1f M input samples per step; reference blocks driven in one large chunk per "
              "thread, channelizer time-sharded and demod tail channel-sharded over %d host threads; FFTW absent -> scalar "
              "float32 mixed-radix FFT stand-in; block construction untimed" % (rows, M, rows * M / 1e6, cores))
    line = {
        "impl": "reference", "metric": "input MS/s, PFB channelizer+DMR demod", "value": value, "unit": "MS/s", "n_gpus": args.gpus,
        "steps": args.steps_ref, "warmup": args.warmup_ref, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "rows_per_step": rows},
        "cpu_baseline": {"value": value, "unit": "MS/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def _dbg(msg):
    if os.environ.get("BENCH_DEBUG"):
        sys.stderr.write("[bench rank %s %.1f] %s\n" % (os.environ.get("RANK", "0"), time.time() % 1000, msg))
        sys.stderr.flush()


def run_ours(args):
    # A time shard keeps ~12 streams busy (front, clock-recovery chain, correlator chain, halo, the chain object's own,
    # one per NCCL communicator), several of them with kernels that spin until a peer arrives.  With the default 8
    # hardware work queues, streams share queues, and a kernel can sit behind a spinning one of another stream: a
    # cross-rank deadlock.  32 is the maximum.  (Must be set before the CUDA context exists.)
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    import numpy as np
    import torch
    import torch.distributed as dist
    from grb200 import chain, lib, sharding, synth_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # a time shard runs its front AFTER its own tail (sharding.TimeShardPlan.front_after_own_tail): no kernel has to
        # co-reside with the clock-recovery kernel, so the chain takes the full-register FFT build (read at chain creation)
        os.environ.setdefault("GRCUDA_CHAIN_NO_OVERLAP", "1")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the product has no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = pin_to_gpu_numa_node(local)
    json_fd = None
    if world > 1:
        # Whatever a library writes to file descriptor 1 from here on (NCCL banners, ...) goes to stderr; the JSON line
        # is written to the original stdout at the end.
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        # Keep stdout to the one JSON line.  With NCCL_DEBUG=VERSION NCCL prints its banner ("NCCL version ...") to
        # stdout and ignores NCCL_DEBUG_FILE (the file is only honoured above that level): drop that level; any more
        # verbose level the caller asked for is kept and sent to stderr.
        if os.environ.get("NCCL_DEBUG", "").strip().upper() == "VERSION":
            del os.environ["NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    R = args.rows
    cfg0 = chain_config(R)
    probe = chain.DmrChain(chain_config(512))
    halo = probe.warmup_rows() if world > 1 else 0   # extra input rows re-processed by each shard
    if world > 1:
        # Load every kernel of the sharded schedule before the first NCCL operation: the first launch of a kernel loads
        # its module, which synchronises the device -- for ever, if a receive kernel is spinning on it for a peer that
        # is itself waiting to load a kernel.
        xp = torch.zeros((probe.history_rows() + 512, M), dtype=torch.complex64, device=dev)
        s0 = torch.cuda.current_stream().cuda_stream
        probe.process_front_device(xp, 512, s0)
        probe.process_tail_mm_device(None, None, s0)
        probe.process_tail_corr_device(None, None, s0)
        torch.cuda.synchronize()
        del xp
    del probe
    cfg = chain_config(R + halo)
    ch = chain.DmrChain(cfg)
    fused_fft = False
    if args.fused_fft:
        # the channelizer output is never written to HBM (the discriminator runs in the last pass of the FFT kernel).
        # Measured SLOWER than the two-kernel default on this machine (profiles/README.md, round 2): the option exists,
        # the bench line is the default path.
        ch.set_keep_channels(False)
        fused_fft = True
    if args.tail_variant is not None:
        ch.set_tail_variant(args.tail_variant)
    if args.fused_correlator:
        ch.set_split_correlator(False)
    if args.split_correlator:
        ch.set_split_correlator(True)
    Th = ch.history_rows()
    # ONE stream for the whole job: the same periodic block on every rank (it tiles seamlessly in time), rank r takes
    # blocks r, r + world, ...  The loop state, and therefore every symbol and sync hit, evolves from block to block.
    x, active = synth_torch.wideband_block(M, R, Th + halo, args.active, 1234, dev)   # [(Th+halo) + R][M]
    plan = sharding.TimeShardPlan(world, rank, R, halo)
    stream = torch.cuda.current_stream().cuda_stream
    if world > 1:
        from grb200 import shardrun
        # Two copies of the input block, used alternately: the NCCL halo exchange of step s+1 writes the head of one
        # while the front of step s still reads the other.
        xbuf = [x, x.clone()]
        sc = shardrun.ShardedChain(ch, plan, dev, R, halo, Th)
        sc.post_halo(0, xbuf)

        def step(s):
            sc.step(s, xbuf, total_steps - 1)   # (total_steps: set below, before the first step)

        def drain():
            sc.drain()
    else:
        ch.set_accumulate_hits(True)
        ch.clear_hits(stream)

        def step(s):
            ch.process_device(x, R, stream)     # on torch's current stream: the timing events live there

        def drain():
            ch.join(stream)

    # Regions: W warm-up steps | K timed steps (the contract's number) | S more steps, >= --sustain-seconds of back-to-back
    # work, timed separately so that the clock sampler sees the steady state (same stream, same chain, continuing).
    def region(first, count, timed):
        _dbg("region %d +%d" % (first, count))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for s in range(first, first + count):
            step(s)
        if world > 1 and first + count < total_steps:
            sc.prepost(first + count)
        drain()          # the tails run on side streams: a region ends when the last one has
        e1.record()
        _dbg("region %d: all launched" % first)
        if world > 1:
            sc.report(4.0)
        torch.cuda.synchronize()
        _dbg("region %d: synchronized" % first)
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # the sustained region's step count must be known before the first step (the very last block sends no state on)
    n_sustain = 0
    if args.sustain_seconds > 0:
        est_ms = {1: 1.9, 2: 2.2, 4: 2.8, 8: 5.2}.get(world, 5.2) * R / 12500.0
        n_sustain = int(min(4000, max(8, args.sustain_seconds * 1e3 / est_ms)))
    main_steps = args.warmup + args.steps
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_warm = 1
    n_e2e = (e2e_warm + e2e_steps) if world > 1 else 0     # N > 1: the end-to-end steps continue the same sharded stream
    total_steps = main_steps + n_sustain + n_e2e
    region(0, args.warmup, False)
    ch.set_profiling(True)
    sampler = ClockSampler(local) if rank == 0 else None
    n0 = lib.launches()
    ms_max = region(args.warmup, args.steps, True)
    launches = lib.launches() - n0
    clocks = sampler.stop() if sampler else None
    prof = ch.profile_read()
    ch.set_profiling(False)
    ms = ms_max

    # ---- final result gather + parity of the whole job against ONE chain running the same stream ----------------------
    # (every sync hit of every block of every rank, as (channel, absolute bit index); outside the timed region)
    parity = None
    _dbg("gathering hits")
    if world > 1:
        allhits = sc.gather_hits()
    else:
        hh, nh_all = ch.read_hits_array(ch.max_hits())
        allhits = torch.stack([torch.from_numpy(hh["channel"].astype("int64")), torch.from_numpy(hh["bit_index"].astype("int64"))], 1).to(dev)
    if rank == 0:
        from grb200 import shardrun as _sr
        csum, cnt = _sr.hit_checksum(allhits)
        parity = {"blocks": world * main_steps, "sync_hits": cnt, "checksum": csum}
        if world > 1 and not args.no_verify:
            ref = chain.DmrChain(cfg)
            if fused_fft:
                ref.set_keep_channels(False)
            ref.set_accumulate_hits(False)
            parts = []
            for b in range(world * main_steps):
                if b == 0:
                    ref.seek_async(-halo, stream)
                    ref.process_front_device(x, halo + R, stream)
                else:
                    ref.process_front_device(x[halo:], R, stream)
                ref.process_tail_device(stream)
                hb, nb = ref.read_hits_array(ref.max_hits())
                parts.append(torch.stack([torch.from_numpy(hb["channel"].astype("int64")),
                                          torch.from_numpy(hb["bit_index"].astype("int64"))], 1))
            refhits = torch.cat(parts).to(dev)
            rsum, rcnt = _sr.hit_checksum(refhits)

            def canon(tt):
                key = tt[:, 1] * 65536 + tt[:, 0]
                return torch.sort(key).values
            same = rcnt == cnt and bool(torch.equal(canon(refhits), canon(allhits)))
            parity.update({"single_chain_sync_hits": rcnt, "single_chain_checksum": rsum, "identical_to_single_chain": same})
            del ref
            assert same, "sharded run differs from the single chain: %r" % (parity,)
        elif world == 1 and fused_fft and not args.no_verify:
            # one GPU: the same blocks through the default two-kernel discriminator path must give the same sync hits
            ref = chain.DmrChain(cfg)
            ref.set_accumulate_hits(True)
            ref.clear_hits(stream)
            for b in range(main_steps):
                ref.process_device(x, R, stream)
            ref.join(stream)
            hb, nb = ref.read_hits_array(ref.max_hits())
            refhits = torch.stack([torch.from_numpy(hb["channel"].astype("int64")), torch.from_numpy(hb["bit_index"].astype("int64"))], 1).to(dev)
            rsum, rcnt = _sr.hit_checksum(refhits)
            key = lambda tt: torch.sort(tt[:, 1] * 65536 + tt[:, 0]).values
            ka, kb = key(refhits), key(allhits)
            common = int(torch.isin(kb, ka).sum().item())
            # The two FFT kernels differ in the last bit of their outputs (like two FFT libraries).  In noise-only stretches
            # the loop's decision feedback turns such a difference into a different symbol COUNT, so later sync words are
            # found at shifted absolute bit indices: the lists agree in number (bounded here) but not index by index over
            # 23 s of signal.  The exact statement is stage isolated and lives in tests/test_gpu_chain.py.
            parity.update({"two_kernel_path_sync_hits": rcnt, "sync_hits_at_identical_bit_index": common,
                           "count_ratio": cnt / max(rcnt, 1)})
            del ref
            assert abs(cnt - rcnt) <= 1e-3 * max(cnt, rcnt), "fused FFT + discriminator path vs two-kernel path: %r" % (parity,)
    _dbg("parity %r" % (parity,))
    # ---- sustained region (clocks under load) -------------------------------------------------------------------------
    sustained = None
    if n_sustain > 0:
        ch.set_accumulate_hits(False)
        sampler2 = ClockSampler(local) if rank == 0 else None
        ms_sus = region(main_steps, n_sustain, True)
        clocks2 = sampler2.stop() if sampler2 else None
        sustained = {"value": world * n_sustain * R * M / (ms_sus * 1e-3) / 1e6, "unit": "MS/s", "steps": n_sustain,
                     "seconds": ms_sus * 1e-3, "ms_per_step": ms_sus / n_sustain, "clocks": clocks2}
    samples_per_step = R * M
    value = world * args.steps * samples_per_step / (ms_max * 1e-3) / 1e6

    _dbg("e2e")
    # ---- end to end: pinned host input, H2D + D2H inside the timed region ----------------------------------------------
    if world == 1:
        # through the host-pointer C ABI
        host = torch.empty((Th + R, M), dtype=torch.complex64, pin_memory=True)
        host.copy_(x[halo: halo + Th + R])
        ch2 = chain.DmrChain(cfg0)
        if fused_fft:
            ch2.set_keep_channels(False)
        ch2.process_host(host.data_ptr(), R)
        ch2.read_hits(16)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        d2h = 0
        for _ in range(e2e_steps):
            ch2.process_host(host.data_ptr(), R)
            _, nh = ch2.read_hits_array(1 << 16)      # the step's result: {channel, bit index} of every sync word found
            d2h += 4 + min(nh, 1 << 16) * 16
        torch.cuda.synchronize()
        te = time.perf_counter() - t0
        e2e_value = e2e_steps * samples_per_step / te / 1e6
        e2e = {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": (Th + R) * M * 8, "d2h_bytes_per_step": d2h // e2e_steps,
               "steps": e2e_steps, "api": "grcuda_dmr_chain_process_host + grcuda_dmr_chain_read_hits (pinned host input)"}
        del ch2
    else:
        # through the SAME sharded schedule as `value`: every rank copies its block (tap history + halo + rows) from its own
        # pinned host buffer (no NCCL halo: the halo rows come with the block), the loop state travels rank to rank as
        # before, and the timed region ends when every rank's sync hits are on rank 0's host
        host = torch.empty((Th + halo + R, M), dtype=torch.complex64, pin_memory=True)
        host.copy_(x)
        copy_ts = torch.cuda.Stream(device=dev)
        copied = [None, None]
        cur = torch.cuda.current_stream(dev)

        def issue_copy(s):
            with torch.cuda.stream(copy_ts):
                if sc.front_ev[s % 2] is not None:
                    copy_ts.wait_event(sc.front_ev[s % 2])      # the front of step s - 2 has read this buffer
                if sc.halo_ev[s % 2] is not None:
                    copy_ts.wait_event(sc.halo_ev[s % 2])       # (a halo exchange of the device-resident regions may still be landing there)
                xbuf[s % 2].copy_(host, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_ts)
                copied[s % 2] = ev

        def e2e_region(first, count, last_region):
            torch.cuda.synchronize()
            dist.barrier()
            ch.clear_hits(stream)
            t0 = time.perf_counter()
            issue_copy(first)
            for s in range(first, first + count):
                if s + 1 < first + count:
                    issue_copy(s + 1)                            # H2D of the next block under this block's kernels
                cur.wait_event(copied[s % 2])
                sc.step(s, xbuf, total_steps - 1, exchange_halo=False)
            if not last_region:
                sc.prepost(first + count)
            sc.drain()
            hits = sc.gather_hits()                              # device -> rank 0 (NCCL) -> host
            nbytes = 0
            if rank == 0:
                hcpu = hits.cpu()
                nbytes = hcpu.numel() * 8
            torch.cuda.synchronize()
            te = time.perf_counter() - t0
            t = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item()), nbytes
        ch.set_accumulate_hits(True)
        first = main_steps + n_sustain
        e2e_region(first, e2e_warm, False)
        te, nbytes = e2e_region(first + e2e_warm, e2e_steps, True)
        e2e_value = world * e2e_steps * samples_per_step / te / 1e6
        e2e = {"value": e2e_value, "unit": "MS/s", "h2d_bytes_per_step": (Th + halo + R) * M * 8,
               "d2h_bytes_per_step": nbytes // max(e2e_steps, 1), "steps": e2e_steps,
               "api": "pinned host block -> H2D -> grcuda_dmr_chain_process_front_device / process_tail_mm_device / "
                      "process_tail_corr_device (time-sharded schedule, loop state by NCCL) -> all sync hits gathered to rank 0's host",
               "numa": numa}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel = the one with the largest device time inside the timed region -------------
    hbm_peak, peak_kind = peaks()
    fp = measured_fp32()
    stage_ms = {k: v[0] for k, v in prof.items()}
    stage_ln = {k: v[1] for k, v in prof.items()}
    busy = sum(stage_ms.values()) or 1.0
    stages = {}
    rows_done = (halo + R) * args.steps
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    inst = {}
    ip = os.path.join(ROOT, "profiles", "instructions.json")   # warp instructions per cfg5 block from the committed ncu capture
    if os.path.exists(ip):
        try:
            inst = json.load(open(ip))
        except Exception:
            inst = {}
    stage_bytes = dict(STAGE_BYTES)
    stage_kernel = dict(STAGE_KERNEL)
    stage_bound = dict(STAGE_BOUND)
    if fused_fft:   # Y is not materialised: FFT 8 B in + 4 B (discriminator) out, matched filter 4 B in + 4 B out
        stage_bytes.update({"pfb_fft": 12.0, "rrc_fir": 8.0})
        stage_kernel.update({"pfb_fft": "fft_demod_kernel (channelizer FFT + quadrature_demod_cf fused)",
                             "rrc_fir": "demod_front_kernel (fir_filter_fff from the discriminator rows)"})
        stage_bound.update({"pfb_fft": "fp32_issue"})
    for k in stage_bytes:
        if stage_ms.get(k, 0) <= 0:
            continue
        bytes_total = stage_bytes[k] * rows_done * M
        gbs = bytes_total / (stage_ms[k] * 1e-3) / 1e9
        st = {"kernel": stage_kernel[k], "ms_per_step": stage_ms[k] / args.steps, "launches_per_step": stage_ln[k] / args.steps,
              "share": stage_ms[k] / busy, "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak, "bound": stage_bound[k]}
        per_launch_ms = stage_ms[k] / max(stage_ln[k], 1)
        if stage_bound[k] == "latency" and fp:
            # cycles one symbol takes vs the dependent-chain floor at the measured instruction latencies
            syms = (halo + R) * SYMBOLS_PER_ROW * (args.steps / max(stage_ln[k], 1))
            cyc = per_launch_ms * 1e-3 * sm_mhz * 1e6 / syms
            floor = MM_CHAIN_OPS * fp["lat_w1"]["fadd"] + MM_CHAIN_LDS * fp["lds_dependent_cycles_w1"]
            st["own_bound"] = {"kind": "latency", "achieved": cyc, "floor": floor, "unit": "cycles/symbol", "frac": floor / cyc,
                               "peak_source": "profiles/r2_fp32_peaks.json"}
        elif stage_bound[k] in ("fp32_issue", "issue") and fp and inst.get(("fused_" if fused_fft else "") + k):
            # warp instructions per second vs the measured issue rate (one instruction per scheduler per cycle)
            wi = inst[("fused_" if fused_fft else "") + k] * (halo + R) / 12500.0
            rate = wi / (per_launch_ms * 1e-3) / 1e9
            st["own_bound"] = {"kind": stage_bound[k], "achieved": rate, "peak": fp["issue_ginst_s"], "unit": "G warp-instructions/s",
                               "frac": rate / fp["issue_ginst_s"], "warp_instructions_per_launch": wi,
                               "peak_source": "profiles/r2_fp32_peaks.json", "count_source": "profiles/instructions.json (ncu)"}
        stages[k] = st
    dom = max(stages, key=lambda k: stages[k]["ms_per_step"])
    d = stages[dom]
    per_launch_bytes = stage_bytes[dom] * rows_done * M / max(stage_ln[dom], 1)
    roofline = {"bound": "hbm", "kernel": dom, "kernel_name": stage_kernel[dom], "achieved": d["algorithmic_GBps"], "peak": hbm_peak,
                "peak_kind": peak_kind, "unit": "GB/s", "frac": d["frac_of_hbm_peak"], "traffic": None,
                "algorithmic_bytes_per_launch": per_launch_bytes,
                "avg_launch_ms": stage_ms[dom] / max(stage_ln[dom], 1), "real_bound": stage_bound[dom],
                "own_bound": d.get("own_bound"), "stages": stages,
                "chain": {"algorithmic_bytes_per_sample": sum(stage_bytes[k] for k in stages),
                          "GBps_over_step": sum(stage_bytes[k] for k in stages) * R * M / (ms_max / args.steps * 1e-3) / 1e9,
                          "frac_of_hbm_peak": sum(stage_bytes[k] for k in stages) * R * M / (ms_max / args.steps * 1e-3) / 1e9 / hbm_peak},
                "note": ROOFLINE_NOTES.get(dom, "")}
    tp = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu capture
    if os.path.exists(tp):
        try:
            roofline["traffic"] = json.load(open(tp)).get(dom)
        except Exception:
            pass

    # ---- CPU baseline beside it (rank 0, N = 1 only; bounded sample) --------------------------------
    cpu = None
    if world == 1 and not args.no_cpu:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import refharness as Rh
            if Rh.available():
                rows_cpu = min(args.cpu_rows, R)
                xs = x[halo + Th: halo + Th + rows_cpu].cpu().numpy().reshape(-1)
                cores = os.cpu_count() or 1
                kw = dict(M=M, pfb_taps=cfg.pfb_taps, quad_gain=cfg.quad_gain, rrc_taps=cfg.rrc_taps, omega=cfg.omega,
                          gain_omega=cfg.gain_omega, mu=cfg.mu, gain_mu=cfg.gain_mu, limit=cfg.omega_relative_limit,
                          slicer_alpha=cfg.slicer_alpha, symbol_map=cfg.symbol_map, access_code=cfg.access_code,
                          threshold=cfg.threshold, x=xs, nthreads=cores, fft_fast=True)
                # bounded sample: repeat the pass over the sample until ~cpu_seconds of CPU work have been timed
                s0 = s1 = 0.0
                reps = nh = 0
                while reps < 1 or (s0 + s1 < args.cpu_seconds and reps < 64):
                    a0, a1, nh, _ = Rh.bench_chain(**kw)
                    s0, s1, reps = s0 + a0, s1 + a1, reps + 1
                cpu = {"value": reps * rows_cpu * M / (s0 + s1) / 1e6, "unit": "MS/s", "cores": cores, "kind": "reference",
                       "sample": "%d passes over the first %d rows (%.1f M samples each) of the GPU workload; the reference's own "
                                 "blocks (oracle/_ref, SSE FIRs) on %d host threads, channelizer time-sharded and demod tail "
                                 "channel-sharded; channelizer %.2f s + demod %.2f s; FFTW absent -> scalar float32 "
                                 "mixed-radix FFT stand-in" % (reps, rows_cpu, rows_cpu * M / 1e6, cores, s0, s1),
                       "sync_hits": nh}
        except Exception as e:  # the baseline must never take the bench down
            cpu = {"value": None, "unit": "MS/s", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}

    # ---- the other BASELINE.json configs and the blocks either side of the path (N = 1 only), each with its own fraction
    extra = None
    if world == 1 and not args.no_extra:
        try:
            del x
            torch.cuda.empty_cache()
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_blocks.py")], stdout=subprocess.PIPE,
                               stderr=subprocess.PIPE, text=True, timeout=240)
            extra = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"failed": r.stderr[-400:]}
        except Exception as e:
            extra = {"failed": repr(e)}

    line = {
        "metric": "input MS/s, PFB channelizer+DMR demod", "value": value, "unit": "MS/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows_per_step_per_gpu": R, "samples_per_step_per_gpu": samples_per_step,
                   "active_channels": int(args.active), "halo_rows": halo, "sharding": "time blocks, block = step*N + rank",
                   "front_vs_own_tail": "after" if plan.front_after_own_tail else "overlapped",
                   "loop_state_handoff": (sc.handoff if world > 1 else None),
                   "channelizer_output": "kept in HBM" if not fused_fft else "not materialised: discriminator fused into the channelizer's FFT kernel (bit identical symbols and hits)",
                   "l2": "input block (%.0f MB) and every intermediate are larger than the 126 MB L2" % (samples_per_step * 8 / 1e6),
                   "host_numa_binding": numa},
        "e2e": e2e,
        "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "parity": parity, "sustained": sustained, "extra": extra,
    }
    if json_fd is not None:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    else:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def _watchdog(seconds):
    """A multi-rank run that stops making progress must fail loudly instead of hanging the box."""
    import signal

    def on_alarm(signum, frame):
        sys.stderr.write("bench.py: no result after %d s, aborting\n" % seconds)
        sys.stderr.flush()
        os._exit(3)
    signal.signal(signal.SIGALRM, on_alarm)
    signal.alarm(seconds)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=12500, help="channel-rate rows per step per GPU (12500 = 1 s of signal)")
    ap.add_argument("--active", type=int, default=800, help="channels carrying DMR bursts (10 % occupancy)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-rows", type=int, default=4096, help="rows of one pass of the CPU baseline sample")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU work to time for the baseline beside the GPU number")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--fused-fft", action="store_true", help="keep_channels = 0: discriminator inside the channelizer's FFT kernel, the channelizer output is not written to HBM")
    ap.add_argument("--no-extra", action="store_true", help="skip the other configs / single blocks (tools/bench_blocks.py) at N = 1")
    ap.add_argument("--sustain-seconds", type=float, default=2.0, help="length of the extra, separately timed steady-state region (0: none)")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the comparison of all sync hits with a single chain")
    ap.add_argument("--tail-variant", type=int, default=None, help="build of the clock-recovery kernel (default: the chain's choice)")
    ap.add_argument("--split-correlator", action="store_true", help="two-kernel tail also underneath the front (single GPU)")
    ap.add_argument("--fused-correlator", action="store_true", help="correlator inside the clock-recovery kernel (round-1 form)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # the reference arm's steps are bounded samples (--cpu-rows) so that K steps finish within minutes
    _watchdog(900)
    args.steps_ref = max(1, args.steps)
    args.warmup_ref = max(0, args.warmup)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
