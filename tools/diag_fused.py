#!/usr/bin/env python3
"""Lab tool: where does the keep_channels = 0 chain differ from the default one (first differing symbol per channel)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gnuradio-3.5.0-dmr_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from grb200 import chain, synth
from test_gpu_chain import make_cfg
M, T, rows = 8000, 16, 1500
rng = np.random.default_rng(M + T)
x, _ = synth.wideband_compose(rng, M, rows, [0, 1, M // 2, M - 1, 1234, 77], noise_sigma=1e-3)
xr = x.reshape(rows, M)
cfg = make_cfg(M, T, max_rows=rows, keep_bytes=False)
res = []
for keep in (True, False):
    ch = chain.DmrChain(cfg)
    if not keep:
        ch.set_keep_channels(False)
    buf = torch.from_numpy(np.concatenate([np.zeros((ch.history_rows(), M), np.complex64), xr])).cuda()
    ch.process_device(buf, rows)
    torch.cuda.synchronize()
    r = ch.fetch()
    res.append((r["counts"].copy(), r["soft"].copy()))
(ca, sa), (cb, sb) = res
print("counts equal", np.array_equal(ca, cb))
mask = np.arange(sa.shape[0])[:, None] < np.minimum(ca, cb)[None, :]
diff = (sa.view(np.uint32) != sb.view(np.uint32)) & mask
print("differing symbols", int(diff.sum()), "of", int(mask.sum()), "channels with a difference", int(diff.any(0).sum()))
first = np.where(diff.any(0), diff.argmax(0), -1)
chs = np.nonzero(first >= 0)[0]
print("channels", chs[:40])
print("first differing symbol index", first[chs][:40])
print("approx row", (first[chs][:40] * 2.604).astype(int))
if len(chs):
    c = chs[0]; k = first[c]
    print("values", sa[k:k + 4, c], sb[k:k + 4, c])
print("channel mod 400 of differing channels:", np.unique(chs % 400)[:50])
