#!/usr/bin/env python3
"""Summarises `ncu --page source --csv` output: stall samples and executed instructions per address range
(e.g. the warp roles of a warp-specialised kernel) and the hottest instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_source_hot.py src.csv [split_hex ...]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]
ia, isrc = h.index('Address'), h.index('Source')
iss, iex = h.index('Warp Stall Sampling (All Samples)'), h.index('Instructions Executed')
splits = sorted(int(x, 16) for x in sys.argv[2:])
base = int(rows[2][ia], 16)
stallcols = [i for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
samples, execd, agg = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
recs = []
for r in rows[2:]:
    if r and r[0] in ('Kernel Name', 'Address'):
        break  # a second view of the same kernel follows: the first one is enough
    if len(r) <= iex:
        continue
    off = int(r[ia], 16) - base
    seg = sum(1 for s in splits if off >= s)
    s, e = int(r[iss]), int(r[iex])
    samples[seg] += s
    execd[seg] += e
    top = sorted([(int(r[i]), h[i][6:]) for i in stallcols if r[i] not in ('', '0')], reverse=True)[:2]
    for i in stallcols:
        if r[i] not in ('', '0'):
            agg[seg][h[i][6:]] += int(r[i])
    recs.append((off, s, e, r[isrc].strip(), top, seg))
tot = sum(samples.values())
print("total samples", tot)
for seg in sorted(samples):
    print("segment %d: samples %d (%.1f%%) warp-instructions %d  top stalls %s" % (seg, samples[seg], 100.0 * samples[seg] / tot, execd[seg], agg[seg].most_common(5)))
for off, s, e, src, top, seg in recs:
    if s >= tot * 0.004:
        print("%5x seg%d %6d %5.1f%% ex=%9d  %-58s %s" % (off, seg, s, 100.0 * s / tot, e, src[:58], top))
