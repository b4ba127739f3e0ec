"""Seeded input generators shared by the CPU and GPU parity tests."""
import json
import os
import random

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dp_golden.json")


def golden():
    with open(GOLDEN) as f:
        return json.load(f)


def enc(s):
    return s.encode("latin1")


class Gen:
    def __init__(self, seed):
        self.rnd = random.Random(seed)

    def rs(self, n, alpha="ACGT"):
        r = self.rnd
        return bytes(ord(r.choice(alpha)) for _ in range(n))

    def mutate(self, s, rate, alpha="ACGTN"):
        r = self.rnd
        out = bytearray()
        for c in s:
            x = r.random()
            if x < rate / 3:
                continue
            if x < 2 * rate / 3:
                out.append(ord(r.choice(alpha)))
            elif x < rate:
                out.append(ord(r.choice(alpha)))
                out.append(c)
            else:
                out.append(c)
        return bytes(out)

    def pair(self, maxn, it, minn=1):
        r = self.rnd
        a = self.rs(r.randint(minn, maxn), "ACGTN" if it % 3 == 0 else "ACGT")
        b = self.mutate(a, r.choice([0, 0.02, 0.1, 0.3])) or b"A"
        if it % 7 == 0:
            b = self.rs(r.randint(minn, maxn))
        if it % 11 == 0:
            b = a
        return a, b

    def borders_case(self):
        r = self.rnd
        p = self.rs(r.randint(1, 60))
        cut = r.randint(0, len(p))
        t = self.mutate(p[:cut], 0.05) + self.rs(r.randint(0, 120)) + self.mutate(p[cut:], 0.05)
        if len(t) < 2:
            t += b"AC"
        return p, t, r.randint(0, 12)

    def gap_case(self):
        r = self.rnd
        ex1, ex2 = self.rs(r.randint(5, 30)), self.rs(r.randint(5, 30))
        intron = b"GT" + self.rs(r.randint(0, 140)) + b"AG"
        est = self.mutate(ex1 + ex2, r.choice([0, 0.05, 0.2])) or b"A"
        return est, ex1 + intron + ex2

    def genome(self, n, n_repeats=4):
        r = self.rnd
        g = bytearray(self.rs(n))
        rep = bytes(g[100:160])
        for _ in range(n_repeats):
            pos = r.randint(200, n - 100)
            g[pos:pos + 60] = rep
        q = r.randint(200, n - 100)
        g[q:q + 10] = b"N" * 10
        return bytes(g)

    def est_from(self, g, it=1):
        r = self.rnd
        parts, pos = [], r.randint(0, max(1, len(g) // 10))
        for _ in range(r.randint(1, 6)):
            ln = r.randint(20, 150)
            parts.append(g[pos:pos + ln])
            pos += ln + r.randint(50, 400)
            if pos >= len(g) - 200:
                break
        e = bytearray(b"".join(parts))
        for _ in range(r.randint(0, 6)):
            e[r.randrange(len(e))] = ord(r.choice("ACGTN"))
        if it % 5 == 0:
            e += b"*" * 20
        if it % 6 == 0:
            e = bytearray(b"#" * 17) + e
        return bytes(e)
