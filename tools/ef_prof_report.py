#!/usr/bin/env python
"""Aggregate ef_prof.out by function: python tools/ef_prof_report.py <binary> <ef_prof.out> [--lines]"""
import subprocess, sys, collections
exe, out = sys.argv[1], sys.argv[2]
by_line = "--lines" in sys.argv
rows = [l.split() for l in open(out) if not l.startswith("lost")]
base = 0
for l in subprocess.run(["readelf", "-lW", exe], capture_output=True, text=True).stdout.splitlines():
    w = l.split()
    if w and w[0] == "LOAD":
        base = int(w[2], 16) - int(w[1], 16); break
# __executable_start is the first LOAD's vaddr
first = None
for l in subprocess.run(["readelf", "-lW", exe], capture_output=True, text=True).stdout.splitlines():
    w = l.split()
    if w and w[0] == "LOAD":
        first = int(w[2], 16); break
addrs = [hex(int(r[0], 16) + first) for r in rows]
res = subprocess.run(["addr2line", "-f", "-e", exe] + addrs, capture_output=True, text=True).stdout.splitlines()
agg = collections.Counter(); agg_libc = collections.Counter()
for i, r in enumerate(rows):
    fn, loc = res[2 * i], res[2 * i + 1].split("/")[-1].split(" ")[0]
    key = f"{fn} {loc}" if by_line else fn
    agg[key] += int(r[1]); agg_libc[key] += int(r[2])
tot = sum(agg.values()) + sum(agg_libc.values())
print(f"samples {tot} (1 ms each)")
for k, _ in sorted(agg.items(), key=lambda kv: -(kv[1] + agg_libc[kv[0]]))[:45]:
    print(f"{100 * (agg[k] + agg_libc[k]) / tot:6.2f}%  self {agg[k]:7d}  in-libc-callee {agg_libc[k]:7d}  {k}")
