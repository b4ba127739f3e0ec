#!/usr/bin/env python
"""Replay a PC_CAPTURE file batch by batch on the GPU and compare EVERY job with the oracle port (developer tool / GPU box).

  python tools/check_capture.py <capture> <genomic.txt>      (genome = est-fact's: N tails stripped, as uploaded)
"""
import ctypes
import sys
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pintron_b200
from pintron_b200.binding import JOB_DTYPE, PC_B_IN_GENOME, PC_OP, PC_RES_INTS
from oracle.binding import Port

cap, gfile = sys.argv[1], sys.argv[2]
lines = open(gfile, "rb").read().split(b"\n")
genome = b"".join(l.rstrip() for l in lines[1:])
genome = genome.strip(b"Nn") if False else genome
# est-fact strips N tails (io-multifasta.c:830): leading / trailing runs of N
g0 = len(genome) - len(genome.lstrip(b"Nn"))
genome = genome.strip(b"Nn")
port = Port()
cu = pintron_b200.Cuda(0)
cu.genome_upload(genome, 15, 0.2)
raw = np.fromfile(cap, dtype=np.uint8)
at, nb, bad, total = 0, 0, 0, 0
names = ["ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX", "SUFCUT", "PRECUT", "LCS", "SEED"]
while at + 12 <= raw.size:
    n = int(raw[at:at + 4].view("<u4")[0]); ab = int(raw[at + 4:at + 12].view("<u8")[0]); at += 12
    if n == 0xffffffff:      # marker written by pc_submit_parts: the next `ab` records ran as one device batch
        continue
    jobs = raw[at:at + n * 44].view(JOB_DTYPE).copy(); at += n * 44
    arena = raw[at:at + ab].copy(); at += ab
    arena = np.concatenate([arena, np.zeros(16, np.uint8)])
    op = jobs["op"]
    size = np.where((op == 0) | (op == 4), jobs["out_cap"].astype(np.int64), np.where(op == 9, 12 * jobs["out_cap"].astype(np.int64), 0))
    size = (size + 3) & ~3
    off = np.concatenate(([0], np.cumsum(size)))
    jobs["out_off"] = off[:-1].astype(np.uint32)
    res, var = cu.run_arrays(arena[:max(ab, 1)], jobs, int(off[-1]))
    ab_bytes = arena.tobytes()
    for q in range(n):
        j, r = jobs[q], res[q]
        a = ab_bytes[j["a_off"]:j["a_off"] + j["a_len"]]
        src = genome if (j["flags"] & PC_B_IN_GENOME) else ab_bytes
        b = src[j["b_off"]:j["b_off"] + j["b_len"]]
        o = int(j["op"]); total += 1
        ok = True
        if o == PC_OP.EDIT:
            ok = r[0] == 0 and r[1] == port.edit(a, b)
        elif o == PC_OP.KBAND:
            ok = r[0] == 0 and (bool(r[1]), int(r[2])) == port.kband(a, b, int(j["p0"]))
        elif o == PC_OP.BORDERS:
            text = src[j["b_off"]:j["b_off"] + j["b_len"] + 2]
            if j["flags"] & 2:
                text = b + b"\0"
            out = (ctypes.c_int * 4)()
            okp = bool(port.lib.po_borders(a, len(a), int(j["p1"]), int(j["p2"]), text, int(j["b_len"]), ctypes.c_uint(int(j["p0"]) & 0xffffffff), out))
            ok = r[0] == 0 and bool(r[1]) == okp and list(r[2:6]) == list(out)
        elif o == PC_OP.GAP:
            ops, pos = port.gap(a, b)
            ok = r[0] == 0 and var[j["out_off"]:j["out_off"] + r[1]].tobytes() == ops and list(r[2:7]) == pos
        elif o == PC_OP.ALIGN:
            s, ops = port.align(a, b)
            ok = r[0] == 0 and r[1] == s and var[j["out_off"]:j["out_off"] + r[2]].tobytes() == ops
        elif o == PC_OP.LCS:
            ok = r[0] == 0 and tuple(r[1:4]) == port.lcs(b, a)
        elif o == PC_OP.AFFIX:
            okp, e_, g_ = port.affix(a, b)
            ok = r[0] == 0 and bool(r[1]) == okp and (not okp or (r[2], r[3]) == (e_, g_))
        if not ok:
            bad += 1
            if bad <= 12:
                print("MISMATCH batch", nb, "job", q, names[o], "a_len", j["a_len"], "b_len", j["b_len"], "p", j["p0"], j["p1"], j["p2"], "flags", j["flags"], "res", list(r))
                print("   a =", a[:80], " b =", b[:80])
    nb += 1
print(f"checked {total} jobs in {nb} batches: {bad} mismatches")
