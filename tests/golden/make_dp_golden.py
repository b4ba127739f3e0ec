#!/usr/bin/env python
"""Generate tests/golden/dp_golden.json from the UNMODIFIED reference routines.

Runs only where oracle/_ref/libref_dp.so exists (the container that has /root/reference):
    make -C oracle ref && python tests/golden/make_dp_golden.py
Every record holds the inputs and the outputs the reference produced, so the fixture pins the oracle
port (tests/test_oracle_port.py) and the CUDA path (tests/test_gpu_parity.py) on any machine.
Alignment rows are stored as the reference printed them ('-' for gaps).
"""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.binding import Ref  # noqa: E402

rnd = random.Random(20261018)


def rs(n, alpha="ACGT"):
    return bytes(ord(rnd.choice(alpha)) for _ in range(n))


def mutate(s, rate, alpha="ACGTN"):
    out = bytearray()
    for c in s:
        r = rnd.random()
        if r < rate / 3:
            continue
        if r < 2 * rate / 3:
            out.append(ord(rnd.choice(alpha)))
        elif r < rate:
            out.append(ord(rnd.choice(alpha)))
            out.append(c)
        else:
            out.append(c)
    return bytes(out)


def pair(maxn, it):
    a = rs(rnd.randint(1, maxn), "ACGTN" if it % 3 == 0 else "ACGT")
    b = mutate(a, rnd.choice([0, 0.02, 0.1, 0.3])) or b"A"
    if it % 7 == 0:
        b = rs(rnd.randint(1, maxn))
    if it % 11 == 0:
        b = a
    return a, b


def main():
    R = Ref()
    d = lambda x: x.decode("latin1")
    G = {k: [] for k in ("align", "edit", "kband", "borders", "gap", "affix", "suffix_cut", "prefix_cut", "lcs",
                         "burset", "seed")}
    for it in range(60):
        a, b = pair(160 if it < 50 else 600, it)
        s, ra, rb = R.align(a, b)
        G["align"].append({"est": d(a), "gen": d(b), "score": s, "est_row": d(ra), "gen_row": d(rb)})
        G["edit"].append({"a": d(a), "b": d(b), "dist": R.edit(a, b)})
        k = rnd.randint(0, 12)
        ok, e = R.kband(a, b, k)
        G["kband"].append({"a": d(a), "b": d(b), "k": k, "ok": ok, "edit": e})
        ed, c1, c2 = R.suffix_cut(a, b)
        G["suffix_cut"].append({"a": d(a), "b": d(b), "out": [ed, c1, c2]})
        ed, c1, c2 = R.prefix_cut(a, b)
        G["prefix_cut"].append({"a": d(a), "b": d(b), "out": [ed, c1, c2]})
        G["affix"].append({"est": d(a), "gen": d(b), "out": list(R.affix(a, b))})
        ln, o1, o2 = R.lcs(a, b)
        G["lcs"].append({"s1": d(a), "s2": d(b), "out": [ln, o1, o2]})
        p = rs(rnd.randint(1, 40))
        cut = rnd.randint(0, len(p))
        t = mutate(p[:cut], 0.05) + rs(rnd.randint(0, 80)) + mutate(p[cut:], 0.05)
        if len(t) < 2:
            t += b"AC"
        me = rnd.randint(0, 10)
        ok, out = R.borders(p, t, me)
        G["borders"].append({"p": d(p), "t": d(t), "max_errs": me, "ok": ok, "out": out})
        ex1, ex2 = rs(rnd.randint(5, 30)), rs(rnd.randint(5, 30))
        intron = b"GT" + rs(rnd.randint(0, 140)) + b"AG"
        est = mutate(ex1 + ex2, rnd.choice([0, 0.05, 0.2])) or b"A"
        gen = ex1 + intron + ex2
        ra, rb, pos = R.gap(est, gen)
        G["gap"].append({"est": d(est), "gen": d(gen), "est_row": d(ra), "gen_row": d(rb), "pos": pos})
    for dn in ("GT", "GC", "AT", "gt", "TT", "NN", "AG", "CT"):
        for ac in ("AG", "AC", "ag", "GG", "TT", "AT", "CA", "NN"):
            G["burset"].append({"donor": dn, "acceptor": ac, "freq": R.burset(dn.encode(), ac.encode())})
    # seeding: a 3 kbp genome with a 5-copy repeat and an N run; ESTs are exon chains with errors
    g = bytearray(rs(3000))
    rep = g[100:160]
    for pos in (500, 900, 1500, 2200):
        g[pos:pos + 60] = rep
    g[1000:1010] = b"N" * 10
    g = bytes(g)
    ix = R.index(g, 15, 0.2)
    seeds = {"genome": d(g), "rate": 0.2, "cases": []}
    for it in range(24):
        parts, pos = [], rnd.randint(0, 300)
        for _ in range(rnd.randint(1, 5)):
            ln = rnd.randint(20, 150)
            parts.append(g[pos:pos + ln])
            pos += ln + rnd.randint(50, 400)
        e = bytearray(b"".join(parts))
        for _ in range(rnd.randint(0, 6)):
            e[rnd.randrange(len(e))] = ord(rnd.choice("ACGTN"))
        if it % 5 == 0:
            e += b"*" * 20
        if it % 6 == 0:
            e = bytearray(b"#" * 17) + e
        e = bytes(e)
        mfl = (15, 16, 20)[it % 3]
        seeds["cases"].append({"est": d(e), "mfl": mfl, "pairings": R.seed(ix, e, mfl)})
    G["seed"] = seeds
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dp_golden.json")
    with open(out, "w") as f:
        json.dump(G, f, separators=(",", ":"))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
