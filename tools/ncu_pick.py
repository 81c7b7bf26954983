#!/usr/bin/env python
"""Print selected metrics per kernel of an ncu report: python tools/ncu_pick.py report.ncu-rep [substr ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
pats = sys.argv[2:] or ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor",
                        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts", "registers_per_thread", "sm__cycles_elapsed.max",
                        "smsp__inst_executed.sum", "issue_active", "lts__t_sector_hit_rate", "l1tex__data_pipe_tma",
                        "smsp__average_warp", "sm__inst_executed_pipe_uniform", "stalled", "dram__throughput", "Grid Size", "Block Size"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
for r in rows[2:]:
    print("=====", r[h.index("ID")], r[h.index("Kernel Name")][:70])
    for i, c in enumerate(h):
        if any(p in c for p in pats) and r[i] not in ("", "0"):
            print(f"  {c:95s} {r[i]:>18s} {units[i]}")
