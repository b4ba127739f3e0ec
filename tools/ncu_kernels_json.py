#!/usr/bin/env python
"""ncu raw page (csv) -> the per-op entry bench.py reads from profiles/r2_ncu_kernels.json: for every op the LONGEST launch of
its kernel (the full-size launch of the merged batch): duration, ALU-pipe %, issue-active %, warps-active %, registers, DRAM
bytes.   python tools/ncu_kernels_json.py <full_raw.csv> <workload> [reads]"""
import csv
import json
import sys

OPS = {"GAP": "k_gap_pairs", "LCS": "k_lcs_len", "BORDERS": "k_borders_packed", "EDIT": "k_myers<(pc_op)2", "KBAND": "k_myers<(pc_op)1",
       "SEED": "k_seed", "ALIGN": "k_warp_per_job<0", "AFFIX": "k_warp_per_job<5"}
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = rows[0]
k = hdr.index("Kernel Name")


def num(r, name):
    try:
        return float(r[hdr.index(name)].replace(",", ""))
    except (ValueError, IndexError):
        return None


units = rows[1]
out = {}
for op, pat in OPS.items():
    best = None
    for r in rows[2:]:
        if pat in r[k].replace(" ", "") or pat in r[k]:
            d = num(r, "gpu__time_duration.sum") or 0
            if best is None or d > best[0]:
                best = (d, r)
    if not best:
        continue
    d, r = best
    scale = {"us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(units[hdr.index("gpu__time_duration.sum")], 1e-6)
    def byt(name):
        v = num(r, name)
        if v is None:
            return None
        return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}.get(units[hdr.index(name)], 1.0)
    rd, wr = byt("dram__bytes_read.sum"), byt("dram__bytes_write.sum")
    out[op] = {"kernel": r[k][:90], "reads": int(sys.argv[3]) if len(sys.argv) > 3 else 20000, "ms": d * scale,
               "alu_pipe_pct": num(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
               "registers": num(r, "launch__registers_per_thread"), "grid": num(r, "launch__grid_size"),
               "warp_instructions": num(r, "smsp__inst_executed.sum"),
               "dram_bytes_per_launch": (rd or 0) + (wr or 0) if rd is not None or wr is not None else None}
print(json.dumps({sys.argv[2]: out}, indent=1))
