#!/usr/bin/env python3
"""Prints the per-kernel summary of an ncu report that profiles/ keeps (and bench.py's `traffic` reads):
usage: ncu -i rep.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv [traffic.json]"""
import csv
import json
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma pipe %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma pipe cycles %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb /issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb /issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait /issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier /issue"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math throttle /issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected /issue"),
]
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
traffic = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    print("== " + name)
    for key, label in WANT:
        if key in h:
            i = h.index(key)
            print("   %-28s %s %s" % (label, r[i], units[i]))
    try:
        rd, wr = float(r[h.index("dram__bytes_read.sum")]), float(r[h.index("dram__bytes_write.sum")])
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[h.index("dram__bytes_read.sum")]]
        scale_w = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[h.index("dram__bytes_write.sum")]]
        short = ("pfb_fir" if "pfb_fir" in name else "pfb_fft" if "fft_fixed" in name else "rrc_fir" if "demod_front" in name
                 else "mm_slicer" if "mm_" in name else name)
        traffic[short] = rd * scale + wr * scale_w
    except Exception:
        pass
if len(sys.argv) > 2:
    json.dump(traffic, open(sys.argv[2], "w"), indent=1)
    print("wrote", sys.argv[2], traffic)
