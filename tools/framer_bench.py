import sys, os, json
sys.path.insert(0, "gnuradio-3.5.0-dmr_b200")
import torch
from grb200 import blocks as B
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1)
M, nbits = 8000, 9600
by = (torch.rand((nbits, M), generator=g, device=dev) < 0.5).to(torch.uint8)
by |= ((torch.rand((nbits, M), generator=g, device=dev) < 1.0 / 600).to(torch.uint8) << 1)
fr = B.framer_sink_1(M, max_msgs=1 << 18, payload_capacity=1 << 26)
junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): fr.work_device(nbits, by, M, 1)
torch.cuda.synchronize()
tot = 0
for _ in range(5):
    junk.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fr.work_device(nbits, by, M, 1); b.record(); torch.cuda.synchronize(); tot += a.elapsed_time(b)
print("framer ms", tot / 5, "msgs", len(fr.messages()))
