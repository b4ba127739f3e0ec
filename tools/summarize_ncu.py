#!/usr/bin/env python
"""Markdown summary of the two ncu exports tools/profile_bench.sh writes:
    python tools/summarize_ncu.py <launches.csv> <full_raw.csv>"""
import collections
import csv
import re
import sys


def short(name):
    name = name.replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    m = re.match(r"(void )?([A-Za-z0-9_:]+(<[^>(]*>)?)", name)
    s = m.group(2) if m else name
    return s.replace("cub::CUB_200802_SM_1000::", "cub::")[:70]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    k, v = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        n = short(r[k])
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += float(r[v].replace(",", "")) / 1e6
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ms / tot >= 0.002:
            print(f"| `{n}` | {c} | {ms:.2f} | {100 * ms / tot:.1f} % |")
    print(f"| all {sum(a[0] for a in agg.values())} launches | | {tot:.1f} | |")


COLS = [("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("launch__waves_per_multiprocessor", "waves/SM"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write")]


def full(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    k = hdr.index("Kernel Name")
    idx = [(hdr.index(c), lab) for c, lab in COLS if c in hdr]
    seen = collections.OrderedDict()
    for r in rows[2:]:
        n = short(r[k])
        # keep the LONGEST launch of each kernel (the full-size one of the step)
        d = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "") or 0)
        if n not in seen or d > seen[n][0]:
            seen[n] = (d, r)
    print("| kernel | " + " | ".join(lab for _, lab in idx) + " |\n|---|" + "---|" * len(idx))
    for n, (_, r) in seen.items():
        cells = []
        for i, _ in idx:
            u = units[i]
            cells.append(f"{r[i]} {u}".strip())
        print(f"| `{n}` | " + " | ".join(cells) + " |")


if __name__ == "__main__":
    print("### Launch list (gpu__time_duration.sum per launch, summed per kernel)\n")
    launches(sys.argv[1])
    print("\n### --set full, the longest launch of each kernel\n")
    full(sys.argv[2])
