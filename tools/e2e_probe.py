#!/usr/bin/env python3
"""Times grcuda_dmr_chain_process_host on a pinned cfg5 block (experiments on the host staging path)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))
import torch
import bench
from grb200 import chain, lib, synth_torch
lib.load(); torch.cuda.set_device(0); dev = torch.device("cuda", 0)
R, M = 12500, bench.M
ch = chain.DmrChain(bench.chain_config(R))
Th = ch.history_rows()
x, _ = synth_torch.wideband_block(M, R, Th, 800, 1234, dev)
host = torch.empty((Th + R, M), dtype=torch.complex64, pin_memory=True); host.copy_(x)
for _ in range(2):
    ch.process_host(host.data_ptr(), R); ch.read_hits(16)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    ch.process_host(host.data_ptr(), R)
t1 = time.perf_counter()
for _ in range(5):
    ch.process_host(host.data_ptr(), R); ch.read_hits_array(1 << 16)
t2 = time.perf_counter()
d = torch.empty_like(x)
torch.cuda.synchronize(); t3 = time.perf_counter()
for _ in range(5):
    d.copy_(host, non_blocking=True)
torch.cuda.synchronize(); t4 = time.perf_counter()
print("nsub=%s process_host %.2f ms  +read_hits %.2f ms  raw H2D %.2f ms" % (os.environ.get("GRCUDA_CHAIN_HOST_SUBBLOCKS", "default"), (t1 - t0) / 5 * 1e3, (t2 - t1) / 5 * 1e3, (t4 - t3) / 5 * 1e3))
