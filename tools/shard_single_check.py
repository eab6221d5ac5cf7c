#!/usr/bin/env python3
"""One GPU, no NCCL: a chain driven the way a time shard drives it (seek to -halo, front on halo + R rows, the tail as
its two kernels with external state buffers, blocks chained through those buffers) against a chain driven as one
continuous stream.  Sync hits, symbol counts and soft symbols must be identical."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))


def main():
    import numpy as np
    import torch
    import bench
    from grb200 import chain, lib, synth_torch
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
    nblocks = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    M = bench.M
    probe = chain.DmrChain(bench.chain_config(512))
    halo, Th = probe.warmup_rows(), probe.history_rows()
    del probe
    cfg = bench.chain_config(rows + halo)
    x, _ = synth_torch.wideband_block(M, rows, Th + halo, 800, 1234, dev)
    stream = torch.cuda.current_stream().cuda_stream
    # sharded style
    a = chain.DmrChain(cfg)
    u8 = dict(dtype=torch.uint8, device=dev)
    mm_buf = [torch.zeros(a.mm_state_bytes(), **u8) for _ in range(2)]
    co_buf = [torch.zeros(a.corr_state_bytes(), **u8) for _ in range(2)]
    got = []
    for b in range(nblocks):
        a.seek_async(b * rows - halo, stream)
        a.process_front_device(x, halo + rows, stream)
        a.process_tail_mm_device(mm_buf[(b + 1) % 2] if b else None, mm_buf[b % 2], stream)
        a.process_tail_corr_device(co_buf[(b + 1) % 2] if b else None, co_buf[b % 2], stream)
        torch.cuda.synchronize()
        print("sharded-style block", b, "done", flush=True)
        res = a.fetch()
        hits, nh = a.read_hits()
        got.append((res["counts"].copy(), res["soft"][:, ::97].copy(), sorted(hits)))
    # continuous
    c = chain.DmrChain(cfg)
    ok = True
    for b in range(nblocks):
        if b == 0:
            c.seek_async(-halo, stream)
            c.process_front_device(x, halo + rows, stream)
        else:
            c.process_front_device(x[halo:], rows, stream)
        c.process_tail_device(stream)
        torch.cuda.synchronize()
        res = c.fetch()
        hits, nh = c.read_hits()
        cnt = res["counts"]
        same = np.array_equal(cnt, got[b][0]) and sorted(hits) == got[b][2]
        m = np.arange(res["soft"].shape[0])[:, None] < cnt[None, ::97]
        same = same and np.array_equal(np.where(m, res["soft"][:, ::97].view(np.uint32), 0), np.where(m, got[b][1].view(np.uint32), 0))
        print("block", b, "hits", nh, "identical" if same else "DIFFERENT", flush=True)
        ok = ok and same
    print("shard-style vs continuous:", "ok" if ok else "FAILED")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
