#!/usr/bin/env python3
"""Multi-GPU parity of the time-sharded chain, one process per GPU over NCCL:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
      tools/shard_parity.py [--numchans 160] [--rows 1500] [--steps 3]

The stream (world x steps blocks of `rows` rows) is generated identically on every rank; rank r processes
blocks r, r + world, ...: it receives the input halo of each block from the rank that owns the previous
block (NCCL isend/irecv), runs the front, receives the loop state (ring), runs the tail, passes the state
on.  Rank 0 then runs the whole stream through ONE chain and compares: sync hits (channel, absolute bit
index), per-channel symbol counts and every soft symbol must be identical, bit for bit."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--numchans", type=int, default=160)
    ap.add_argument("--taps-per-branch", type=int, default=16)
    ap.add_argument("--rows", type=int, default=1500)
    ap.add_argument("--steps", type=int, default=3)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from grb200 import chain, firdes, lib, sharding, synth

    def log(msg):
        print("[rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()
    M, T, R = args.numchans, args.taps_per_branch, args.rows
    nblocks = world * args.steps
    rng = np.random.default_rng(99)
    active = sorted(set(int(c) for c in rng.choice(M, size=min(24, max(4, M // 10)), replace=False)))  # host-side synthesis is slow
    x, _ = synth.wideband_compose(rng, M, nblocks * R, active, noise_sigma=3e-3)
    xr = x.reshape(nblocks * R, M)
    taps = firdes.low_pass_2(1.0, M * 12500.0, 5500.0, 1500.0, 60.0, firdes.WIN_BLACKMAN_hARRIS)
    c = len(taps) // 2
    taps = (taps[c - M * T // 2: c - M * T // 2 + M * T] * M).astype(np.float32)

    def make(max_rows):
        return chain.DmrChain(chain.DmrChainConfig(M, taps, max_rows_per_block=max_rows))

    probe = make(512)
    halo, Th = probe.warmup_rows(), probe.history_rows()
    del probe
    ch = make(R + halo)
    plan = sharding.TimeShardPlan(world, rank, R, halo)
    ring = sharding.RingExchanger(plan)
    state = torch.zeros(ch.state_bytes(), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    H = Th + halo                                     # rows in front of a block that the front stage re-processes
    recv_buf = torch.zeros((H, M), dtype=torch.complex64, device=dev)
    carried = torch.zeros((H, M), dtype=torch.complex64, device=dev)   # rank 0: halo received one step earlier
    mine = {}
    last = args.steps - 1
    for s in range(args.steps):
        b = plan.block_index(s)
        blk = torch.from_numpy(np.ascontiguousarray(xr[b * R:(b + 1) * R])).to(dev)
        # halo = the last H rows BEFORE this block: they are the tail of block b-1, owned by the left neighbour
        # (same step for rank > 0; the previous step, last rank, for rank 0)
        log("step %d block %d: halo exchange" % (s, b))
        works = ring.exchange_halo(blk[R - H:].contiguous(), recv_buf, s)
        ring.wait_all(works)
        log("step %d: front" % s)
        if world == 1:
            halo_rows = carried
            nxt = blk[R - H:].clone()
        elif rank == 0:
            halo_rows = carried.clone()
            nxt = recv_buf.clone()
        else:
            halo_rows, nxt = recv_buf, None
        buf = torch.cat([halo_rows, blk])             # [Th + halo + R][M]
        if b == 0:
            # stream start: zero history, no warm-up rows to re-process
            buf0 = torch.cat([torch.zeros((Th, M), dtype=torch.complex64, device=dev), blk])
            ch.seek_async(0, stream)
            ch.process_front_device(buf0, R, stream)
        else:
            ch.seek_async(b * R - halo, stream)
            ch.process_front_device(buf, halo + R, stream)
        log("step %d: waiting for the loop state" % s)
        if ring.recv_state(state, s):
            ch.import_state(state, stream)
        log("step %d: tail" % s)
        ch.process_tail_device(stream)
        ch.export_state(state, stream)
        # Read the results back (device-wide synchronisation) BEFORE the state goes out: a device sync while the
        # NCCL send is pending would wait for the right neighbour's recv, which it only posts after the NEXT halo
        # exchange -- in which this rank would then be missing.
        res = ch.fetch()
        hits, nh = ch.read_hits()
        assert nh == len(hits)
        ring.send_state(state, s, last)
        log("step %d: state sent" % s)
        mine[b] = (res["counts"].copy(), [res["soft"][:res["counts"][cc], cc].copy() for cc in active[:6]], sorted(hits))
        if nxt is not None:
            carried = nxt
    ring.finish()
    # gather everything on rank 0 (through files: the ranks share the node)
    import pickle
    import tempfile
    tag = os.environ.get("MASTER_PORT", "0")
    path = os.path.join(tempfile.gettempdir(), "shard_parity_%s_rank%%d.pkl" % tag)
    with open(path % rank, "wb") as f:
        pickle.dump(mine, f)
    log("results written")
    if world > 1:
        dist.barrier()
    gathered = [pickle.load(open(path % r, "rb")) for r in range(world)] if rank == 0 else None
    ok = True
    if rank == 0:
        allb = {}
        for g in gathered:
            allb.update(g)
        assert sorted(allb) == list(range(nblocks))
        # sequential run of the whole stream on one chain
        seq = make(R)
        buf = torch.from_numpy(np.concatenate([np.zeros((Th, M), np.complex64), xr])).to(dev)
        tot_hits, nsym = 0, 0
        for b in range(nblocks):
            seq.process_device(buf[b * R:], R)
            res = seq.fetch()
            hits, _ = seq.read_hits()
            counts, softs, h = allb[b]
            if b == 0:
                pass
            # a shard's first `halo` rows are re-processed warm-up rows: its tail produces the same symbols
            # because it starts from the imported state at the same absolute position
            if not np.array_equal(counts, res["counts"]):
                print("block %d: symbol counts differ" % b)
                ok = False
            for k, cc in enumerate(active[:6]):
                if not np.array_equal(softs[k], res["soft"][:res["counts"][cc], cc]):
                    print("block %d channel %d: soft symbols differ" % (b, cc))
                    ok = False
            if sorted(hits) != h:
                print("block %d: hits differ (%d vs %d)" % (b, len(hits), len(h)))
                ok = False
            tot_hits += len(h)
            nsym += int(res["counts"].sum())
        print("shard parity %s: world %d, %d blocks x %d rows x %d channels, %d symbols, %d sync hits"
              % ("ok" if ok else "FAILED", world, nblocks, R, M, nsym, tot_hits))
    if world > 1:
        flag = torch.tensor([1 if ok else 0], device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
        dist.destroy_process_group()
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
