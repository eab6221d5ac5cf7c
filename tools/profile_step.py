#!/usr/bin/env python3
"""Short, deterministic run of the flagship chain for ncu: `--steps` passes of the cfg5 block
(12 500 rows x 8000 channels by default) through grcuda_dmr_chain_process_device, nothing else.
Each step launches the same kernels in the same order, so `ncu -s <launches of the warm-up steps>
-c <launches of one step>` captures exactly one steady-state step.

  python tools/profile_step.py --steps 3           # plain run (must exit 0 before any ncu run)
  ncu ... python tools/profile_step.py --steps 3   # see /opt/skills/guides/B200_PROFILING.md
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--rows", type=int, default=12500)
    ap.add_argument("--active", type=int, default=800)
    args = ap.parse_args()
    import torch
    import bench
    from grb200 import chain, lib, synth_torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    lib.load()
    ch = chain.DmrChain(bench.chain_config(args.rows))
    Th = ch.history_rows()
    x, _ = synth_torch.wideband_block(bench.M, args.rows, Th, args.active, 1234, dev)
    stream = torch.cuda.current_stream().cuda_stream
    torch.cuda.synchronize()
    n0 = lib.launches()
    for _ in range(args.steps):
        ch.process_device(x, args.rows, stream)
    torch.cuda.synchronize()
    _, nh = ch.read_hits(16)
    print("steps %d, kernel launches per step %d, sync hits in the last step %d"
          % (args.steps, (lib.launches() - n0) // args.steps, nh))


if __name__ == "__main__":
    main()
