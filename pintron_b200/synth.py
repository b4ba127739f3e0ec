"""Synthetic est-fact inputs of the shapes BASELINE.json names; fixed seeds, no datasets.

genomic.txt / ests.txt text in the reference's input format (one FASTA genome record `>chrN:start:end:strand`,
multi-FASTA ESTs with `/gb=` and `/clone_end=` header fields, io-multifasta.c:279-504) plus the simulated exon
structure of every EST (used only to derive DP job shapes for the device-path bench, never as a truth for parity).
"""
import re

import numpy as np

COMP = bytes.maketrans(b"ACGTNacgtn", b"TGCANtgcan")
_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_LINES = re.compile(rb".{1,70}", re.S)          # FASTA body lines of 70 columns

CONFIGS = {
    # name: genome nt, genes, exons/gene, exon len range, intron len range, reads, read len range, mRNA fraction
    "C3": dict(genome=200_000, genes=1, exons=60, exon_len=(20, 90), intron_len=(80, 6000), reads=100_000,
               read_len=(300, 800), err=0.01, seed=1003),
    "C4": dict(genome=2_000_000, genes=8, exons=40, exon_len=(25, 300), intron_len=(80, 60000), reads=1_000_000,
               read_len=(300, 800), err=0.01, seed=1004, mrna_frac=0.2, mrna_len=(1000, 6000)),
    "C5": dict(genome=320_000, genes=1, exons=200, exon_len=(100, 600), intron_len=(80, 2000), reads=500_000,
               read_len=(10_000, 100_000), err=0.005, seed=1005, long_exons=(4, 5000, 17000)),
    # reduced C4 / C5 shapes for parity tests (multi-gene locus with long 3' UTR exons and mRNAs; titin-like long exons)
    "C4mini": dict(genome=400_000, genes=3, exons=25, exon_len=(25, 300), intron_len=(80, 15000), reads=400,
                   read_len=(300, 800), err=0.01, seed=1104, mrna_frac=0.3, mrna_len=(1000, 5000), utr_len=(500, 3000)),
    "C5mini": dict(genome=300_000, genes=1, exons=50, exon_len=(100, 600), intron_len=(80, 4000), reads=16,
                   read_len=(8000, 20000), err=0.005, seed=1105, long_exons=(4, 3000, 9000)),
    "tiny": dict(genome=20_000, genes=1, exons=12, exon_len=(25, 160), intron_len=(80, 1500), reads=64,
                 read_len=(200, 500), err=0.01, seed=7),
}


def _rand_seq(rng, n):
    return np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)]


def _log_uniform(rng, lo, hi):
    return int(round(np.exp(rng.uniform(np.log(lo), np.log(hi)))))


class Synth:
    def __init__(self, name="C3", reads=None, seed=None):
        cfg = dict(CONFIGS[name])
        if reads is not None:
            cfg["reads"] = reads
        if seed is not None:
            cfg["seed"] = seed
        self.cfg, self.name = cfg, name
        rng = self.rng = np.random.default_rng(cfg["seed"])
        G = cfg["genome"]
        g = _rand_seq(rng, G).copy()
        # ~1 % planted low-complexity / tandem repeats
        for _ in range(max(1, G // 20000)):
            unit = _rand_seq(rng, int(rng.integers(1, 5)))
            ln = int(rng.integers(60, 200))
            pos = int(rng.integers(0, G - ln))
            g[pos:pos + ln] = np.resize(unit, ln)
        # gene models
        self.genes = []
        span = G // cfg["genes"]
        for gi in range(cfg["genes"]):
            for _attempt in range(200):
                ex = [_log_uniform(rng, *cfg["exon_len"]) for _ in range(cfg["exons"])]
                if cfg.get("utr_len"):
                    ex[-1] = int(rng.integers(*cfg["utr_len"]))
                if cfg.get("long_exons"):
                    k, lo_, hi_ = cfg["long_exons"]
                    for q in rng.choice(cfg["exons"], size=k, replace=False):
                        ex[int(q)] = int(rng.integers(lo_, hi_))
                it = [_log_uniform(rng, *cfg["intron_len"]) for _ in range(cfg["exons"] - 1)]
                if sum(ex) + sum(it) < span - 2000:
                    break
            else:
                scale = (span - 2000 - sum(ex)) / max(1, sum(it))
                it = [max(60, int(x * scale)) for x in it]
            pos = gi * span + int(rng.integers(500, max(501, span - (sum(ex) + sum(it)) - 500)))
            exons = []
            for k, el in enumerate(ex):
                exons.append((pos, pos + el))
                pos += el
                if k < len(it):
                    r = rng.random()
                    don, acc = (b"GT", b"AG") if r < 0.94 else ((b"GC", b"AG") if r < 0.99 else (b"AT", b"AC"))
                    g[pos:pos + 2] = np.frombuffer(don, dtype=np.uint8)
                    g[pos + it[k] - 2:pos + it[k]] = np.frombuffer(acc, dtype=np.uint8)
                    pos += it[k]
            self.genes.append(exons)
        self.genome = g.tobytes()
        self.transcripts = []
        for exons in self.genes:
            t = b"".join(self.genome[a:b] for a, b in exons)
            bounds = np.cumsum([0] + [b - a for a, b in exons])
            self.transcripts.append((t, bounds, exons))

    def genome_fasta(self):
        G = len(self.genome)
        lines = [f">chr1:1000001:{1000000 + G}:1".encode()]
        lines += [self.genome[i:i + 70] for i in range(0, G, 70)]
        return b"\n".join(lines) + b"\n"

    def _mutate(self, s, err):
        """Per-base errors at rate err: 60 % substitutions, 20 % insertions (a random base before the original one),
        20 % deletions.  Vectorised: one numpy pass per read."""
        rng = self.rng
        n = len(s)
        r = rng.random(n)
        hit = r < err
        if not hit.any():
            return s
        src = np.frombuffer(s, dtype=np.uint8)
        sub, ins, dele = hit & (r < 0.6 * err), hit & (r >= 0.6 * err) & (r < 0.8 * err), hit & (r >= 0.8 * err)
        cnt = np.ones(n, dtype=np.int64)
        cnt[ins] = 2
        cnt[dele] = 0
        off = np.concatenate(([0], np.cumsum(cnt)))
        out = np.empty(int(off[-1]), dtype=np.uint8)
        keep = ~dele
        out[off[1:][keep] - 1] = src[keep]                      # the original base is the last byte of its slot
        rnd = _ACGT[rng.integers(0, 4, size=n)]
        out[off[:-1][sub]] = rnd[sub]
        out[off[:-1][ins]] = rnd[ins]
        return out.tobytes()

    def reads(self, start=0, count=None):
        """Yield (header, sequence, exon_pieces, forward_sequence); exon_pieces = [(genome_start, genome_end)] of the
        error-free read, forward_sequence = the read before the optional reverse-complement."""
        cfg, rng = self.cfg, np.random.default_rng(self.cfg["seed"] * 7919 + start)
        self.rng = rng
        count = cfg["reads"] - start if count is None else count
        for idx in range(start, start + count):
            t, bounds, exons = self.transcripts[int(rng.integers(0, len(self.transcripts)))]
            lo, hi = cfg["read_len"]
            if cfg.get("mrna_frac") and rng.random() < cfg["mrna_frac"]:
                lo, hi = cfg["mrna_len"]
            ln = min(int(rng.integers(lo, hi + 1)), len(t))
            s0 = int(rng.integers(0, len(t) - ln + 1))
            pieces = []
            k0 = int(np.searchsorted(bounds, s0, side="right")) - 1
            k1 = int(np.searchsorted(bounds, s0 + ln, side="left"))
            for k in range(k0, k1):                      # the exons the read overlaps
                a = exons[k][0]
                bk = int(bounds[k])
                x0, x1 = max(s0, bk), min(s0 + ln, int(bounds[k + 1]))
                if x0 < x1:
                    pieces.append((a + x0 - bk, a + x1 - bk))
            seq = self._mutate(t[s0:s0 + ln], cfg["err"])
            if rng.random() < 0.002 * 50:          # ~0.2 % N overall, concentrated in 10 % of the reads
                sa = bytearray(seq)
                for _ in range(max(1, len(sa) // 50)):
                    sa[int(rng.integers(0, len(sa)))] = ord("N")
                seq = bytes(sa)
            if rng.random() < 0.3:
                seq += b"A" * int(rng.integers(15, 41))
            end, fwd = "3'", seq
            if rng.random() < 0.5:
                seq = seq.translate(COMP)[::-1]
                end = "5'"
            yield f">/gb=SYN{idx:07d}/clone_end={end}".encode(), seq, pieces, fwd

    def ests_fasta(self, start=0, count=None):
        out = []
        for h, s, _, _ in self.reads(start, count):
            out.append(h)
            out += _LINES.findall(s)
        return b"\n".join(out) + b"\n"


def ests_fasta_parallel(name, total, start, count, chunk=4000, procs=None):
    """ests.txt bytes for reads [start, start+count) of Synth(name, reads=total), generated by child processes
    (`python -m pintron_b200.synth`, so it is safe next to an initialised CUDA context) in fixed chunks of `chunk` reads;
    each chunk seeds its own generator from its first read index, so the bytes depend on `chunk` but not on the number
    of processes."""
    import os
    import subprocess
    import sys
    import tempfile
    tasks = [(s, min(chunk, start + count - s)) for s in range(start, start + count, chunk)]
    if len(tasks) <= 1:
        return b"".join(Synth(name, reads=total).ests_fasta(s, c) for s, c in tasks)
    procs = procs or max(1, min(len(tasks), (os.cpu_count() or 2) - 1))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = []
    with tempfile.TemporaryDirectory(prefix="pintron_synth_") as tmp:
        running, nxt, files = [], 0, [os.path.join(tmp, f"c{i}.fa") for i in range(len(tasks))]
        while nxt < len(tasks) or running:
            while nxt < len(tasks) and len(running) < procs:
                s, c = tasks[nxt]
                running.append(subprocess.Popen([sys.executable, "-m", "pintron_b200.synth", name, str(total), str(s), str(c), files[nxt]], cwd=root))
                nxt += 1
            p = running.pop(0)
            if p.wait() != 0:
                raise RuntimeError("synthetic EST generation failed")
        for f in files:
            with open(f, "rb") as fh:
                out.append(fh.read())
    return b"".join(out)


if __name__ == "__main__":
    import sys
    _name, _total, _start, _count, _path = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    with open(_path, "wb") as _fh:
        _fh.write(Synth(_name, reads=_total).ests_fasta(_start, _count))
