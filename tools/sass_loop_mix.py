#!/usr/bin/env python3
"""Instruction mix of the sweep loop of k_gap_pairs<LANES, MINB> in a compiled object (cuobjdump -sass): from the loop
head (the SHFL.UP of the wavefront hand-over) to the direction store (STG.E.128).   tools/sass_loop_mix.py k_gap.o [LANES]"""
import re, subprocess, sys
from collections import Counter
obj, lanes = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "16")
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
fn = [m.start() for m in re.finditer(r"Function : ", sass)]
for a, b in zip(fn, fn[1:] + [len(sass)]):
    head = sass[a:a + 200]
    if f"ILi{lanes}ELi4EE" in head:
        body = sass[a:b]
        break
pairs = [(int(m.group(1), 16), m.group(2)) for m in re.finditer(r"/\*([0-9a-f]{4,5})\*/\s+(.*?);", body)]
ins = [t for _, t in pairs]
st = next(i for i, t in enumerate(ins) if "STG.E.128" in t)
lo = max(i for i, t in enumerate(ins[:st]) if "SHFL.UP" in t) - 2
# the fast path of one step: up to the store, then from the store to the branch that skips the (n, m) capture block
hi = st + 1
while not re.match(r"@!?P\d BRA", ins[hi]):
    hi += 1
c = Counter()
for t in ins[lo:hi + 1]:
    t = re.sub(r"^@!?U?P\d+\s+", "@P ", t)
    w = t.split()
    c[(w[0] + " " + w[1]) if w[0] == "@P" else w[0]] += 1
print("loop instructions (head .. store .. branch over the capture block):", hi + 1 - lo, " to the store:", st + 1 - lo)
for k, v in c.most_common(14):
    print(f"  {v:4d} {k}")
regs = re.search(r"REG:(\d+)", subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout)
print(subprocess.run(["cuobjdump", "-res-usage", obj], capture_output=True, text=True).stdout.count("REG:"), "functions; first REG:", regs.group(1) if regs else "?")
