#!/usr/bin/env python3
"""Experiment driver (GPU box): times the PFB channelizer's two kernels under different chunk sizes and
FFT plan variants.  Usage: python tools/exp_pfb.py [rows]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200"))
import numpy as np
import torch
from grb200 import blocks

M, T = 8000, 16
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
x = torch.randn((T + rows, M, 2), device="cuda").view(torch.float32)
x = torch.view_as_complex(x.reshape(T + rows, M, 2))
y = torch.empty((rows, M), dtype=torch.complex64, device="cuda")
taps = (np.random.default_rng(0).standard_normal(M * T) / M).astype(np.float32)
import ctypes as C
for variant in os.environ.get("VARIANTS", "0,1,2,3").split(","):
    os.environ["GRCUDA_FFT_VARIANT"] = variant
    pfb = blocks.pfb_channelizer_ccf(M, taps)
    L = pfb.L
    L.grcuda_pfb_channelizer_ccf_set_profiling(pfb.h, 1)
    for mb in os.environ.get("CHUNKS", "16,24,48,96,100000").split(","):
        os.environ["GRCUDA_PFB_CHUNK_MB"] = mb
        for _ in range(2):
            pfb.work_device(rows, x, y)
        torch.cuda.synchronize()
        ms = (C.c_float * 2)(); ln = (C.c_int * 2)()
        L.grcuda_pfb_channelizer_ccf_profile_read(pfb.h, ms, ln)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            pfb.work_device(rows, x, y)
        e1.record()
        torch.cuda.synchronize()
        L.grcuda_pfb_channelizer_ccf_profile_read(pfb.h, ms, ln)
        tot = e0.elapsed_time(e1) / reps
        gb = rows * M * 16 / 1e9
        print("fft_variant=%s chunk_mb=%s total %.3f ms (%.1f GS/s)  fir %.3f ms (%.0f GB/s, %d launches)  fft %.3f ms (%.0f GB/s)"
              % (variant, mb, tot, rows * M / tot / 1e6, ms[0] / reps, gb / (ms[0] / reps) * 1e3, ln[0] // reps, ms[1] / reps,
                 gb / (ms[1] / reps) * 1e3), flush=True)
    del pfb
