"""GPU (-m gpu): the CUDA path, called through the C ABI, against (1) the golden vectors the UNMODIFIED reference
produced, (2) the oracle port on seeded inputs, (3) size-independent properties at sizes the oracle cannot reach.
Bit-exact everywhere: this path is integer / byte work."""
import numpy as np
import pytest

import pintron_b200
from pintron_b200 import Batch, PC_OP
from oracle.binding import Port, ops_to_rows
from util_cases import Gen, enc, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cu():
    c = pintron_b200.Cuda(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def port():
    return Port()


@pytest.fixture(scope="module")
def gold():
    return golden()


def test_native_library_is_loaded(cu):
    before = cu.launch_count()
    assert cu.edit_distance(b"ACGT", b"AGT") == 1
    assert cu.launch_count() > before


def test_align_golden(cu, gold):
    b = Batch()
    for c in gold["align"]:
        b.add(PC_OP.ALIGN, enc(c["est"]), enc(c["gen"]))
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for c, r, j in zip(gold["align"], res, jobs):
        assert r[0] == 0 and r[1] == c["score"]
        ops = var[j["out_off"]:j["out_off"] + r[2]].tobytes()
        assert ops_to_rows(ops, enc(c["est"]), enc(c["gen"])) == (enc(c["est_row"]), enc(c["gen_row"]))


def test_scalar_ops_golden(cu, gold):
    b = Batch()
    exp = []
    for c in gold["edit"]:
        b.add(PC_OP.EDIT, enc(c["a"]), enc(c["b"])); exp.append([c["dist"]])
    for c in gold["kband"]:
        b.add(PC_OP.KBAND, enc(c["a"]), enc(c["b"]), p0=c["k"]); exp.append([int(c["ok"]), c["edit"]])
    for c in gold["borders"]:
        b.add(PC_OP.BORDERS, enc(c["p"]), enc(c["t"]), p0=c["max_errs"], p1=0, p2=len(c["p"]))
        exp.append([int(c["ok"])] + c["out"])
    for c in gold["affix"]:
        b.add(PC_OP.AFFIX, enc(c["est"]), enc(c["gen"])); exp.append([int(c["out"][0])] + c["out"][1:])
    for c in gold["suffix_cut"]:
        b.add(PC_OP.SUFCUT, enc(c["a"]), enc(c["b"])); exp.append(c["out"])
    for c in gold["prefix_cut"]:
        b.add(PC_OP.PRECUT, enc(c["a"]), enc(c["b"])); exp.append(c["out"])
    for c in gold["lcs"]:
        b.add(PC_OP.LCS, enc(c["s2"]), enc(c["s1"])); exp.append(c["out"])
    res, _ = cu.run(b)
    for i, (r, e) in enumerate(zip(res, exp)):
        assert r[0] == 0, (i, r)
        assert list(r[1:1 + len(e)]) == e, (i, b.jobs[i][0], list(r), e)


def test_gap_golden(cu, gold):
    b = Batch()
    for c in gold["gap"]:
        b.add(PC_OP.GAP, enc(c["est"]), enc(c["gen"]))
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for c, r, j in zip(gold["gap"], res, jobs):
        assert r[0] == 0
        ops = var[j["out_off"]:j["out_off"] + r[1]].tobytes()
        assert ops_to_rows(ops, enc(c["est"]), enc(c["gen"])) == (enc(c["est_row"]), enc(c["gen_row"]))
        assert list(r[2:7]) == c["pos"]


def test_seed_golden(cu, gold):
    s = gold["seed"]
    cu.genome_upload(enc(s["genome"]), 15, s["rate"])
    b = Batch()
    for c in s["cases"]:
        b.add(PC_OP.SEED, enc(c["est"]), p0=c["mfl"], out_cap=4096)
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for c, r, j in zip(s["cases"], res, jobs):
        assert r[0] == 0
        tri = var[j["out_off"]:j["out_off"] + 12 * r[1]].view(np.int32).reshape(-1, 3)
        assert [list(map(int, x)) for x in tri] == c["pairings"]


def test_mixed_batch_vs_port(cu, port):
    """One heterogeneous batch of ~3000 jobs, every op, seeded; each result equals the oracle's."""
    g = Gen(4242)
    genome = g.genome(6000)
    cu.genome_upload(genome, 15, 0.2)
    b, chk = Batch(), []
    for it in range(300):
        a, c = g.pair(220, it)
        b.add(PC_OP.ALIGN, a, c); chk.append(("align", a, c))
        b.add(PC_OP.EDIT, a, c); chk.append(("edit", a, c))
        k = g.rnd.randint(0, 14)
        b.add(PC_OP.KBAND, a, c, p0=k); chk.append(("kband", a, c, k))
        b.add(PC_OP.SUFCUT, a, c); chk.append(("sc", a, c))
        b.add(PC_OP.PRECUT, a, c); chk.append(("pc", a, c))
        b.add(PC_OP.AFFIX, a, c); chk.append(("affix", a, c))
        b.add(PC_OP.LCS, c, a); chk.append(("lcs", a, c))
        p, t, me = g.borders_case()
        lo = g.rnd.randint(0, len(p)); hi = g.rnd.randint(lo, len(p))
        b.add(PC_OP.BORDERS, p, t, p0=me, p1=lo, p2=hi); chk.append(("borders", p, t, me, lo, hi))
        est, gen = g.gap_case()
        b.add(PC_OP.GAP, est, gen); chk.append(("gap", est, gen))
        e = g.est_from(genome, it)
        mfl = (15, 16, 21)[it % 3]
        b.add(PC_OP.SEED, e, p0=mfl, out_cap=2048); chk.append(("seed", e, mfl))
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for r, j, c in zip(res, jobs, chk):
        assert r[0] == 0, (c[0], list(r))
        kind = c[0]
        if kind == "align":
            s, ops = port.align(c[1], c[2])
            assert r[1] == s and var[j["out_off"]:j["out_off"] + r[2]].tobytes() == ops
        elif kind == "edit":
            assert r[1] == port.edit(c[1], c[2])
        elif kind == "kband":
            assert (bool(r[1]), int(r[2])) == port.kband(c[1], c[2], c[3])
        elif kind == "sc":
            assert tuple(r[1:4]) == port.suffix_cut(c[1], c[2])
        elif kind == "pc":
            assert tuple(r[1:4]) == port.prefix_cut(c[1], c[2])
        elif kind == "affix":
            ok, e_, g_ = port.affix(c[1], c[2])
            assert bool(r[1]) == ok and (not ok or (r[2], r[3]) == (e_, g_))
        elif kind == "lcs":
            assert tuple(r[1:4]) == port.lcs(c[1], c[2])
        elif kind == "borders":
            ok, out = port.borders(c[1], c[2], c[3], c[4], c[5])
            assert bool(r[1]) == ok and list(r[2:6]) == out
        elif kind == "gap":
            ops, pos = port.gap(c[1], c[2])
            assert var[j["out_off"]:j["out_off"] + r[1]].tobytes() == ops and list(r[2:7]) == pos
        elif kind == "seed":
            tri = var[j["out_off"]:j["out_off"] + 12 * r[1]].view(np.int32).reshape(-1, 3)
            assert [tuple(map(int, x)) for x in tri] == port.seed(genome, c[1], c[2], 0.2)


def test_edge_cases(cu, port):
    assert cu.compute_alignment(b"", b"ACG") == port.align(b"", b"ACG")
    assert cu.compute_alignment(b"ACG", b"") == port.align(b"ACG", b"")
    assert cu.compute_alignment(b"ANG", b"ACG") == port.align(b"ANG", b"ACG")
    assert cu.edit_distance(b"", b"") == 0
    assert cu.K_band_edit_distance(b"ACGT", b"ACGT", 0) == (True, 0)
    assert cu.K_band_edit_distance(b"ACGT", b"ACGA", 0) == (False, 1)
    assert cu.K_band_edit_distance(b"ACGTACGTAC", b"ACG", 2) == (False, 7)
    assert cu.find_longest_common_factor_dp(b"", b"ACG") == (0, 0, 0)
    assert cu.find_longest_common_factor_dp(b"TTANGTT", b"GACGA") == (3, 2, 1)
    assert cu.find_longest_affix(b"", b"ACG") == (False, 0, 0)
    assert cu.general_refine_borders(b"", b"ACGTAC", 3) == port.borders(b"", b"ACGTAC", 3)
    assert cu.compute_gap_alignment(b"A", b"C") == port.gap(b"A", b"C")
    res, _ = cu.run(Batch())
    assert res.shape == (0, 8)


def test_long_alignment_band_and_pool_growth(cu, port):
    """mRNA-sized exons: the banded kernel must reproduce the full-matrix oracle, including when the batch needs more
    scratch than the initial pool (pc_stream_sync retries PC_E_POOL jobs)."""
    g = Gen(99)
    b, pairs = Batch(), []
    for it in range(24):
        a = g.rs(g.rnd.randint(1500, 4000))
        c = g.mutate(a, (0.01, 0.03, 0.1)[it % 3], alpha="ACGT")
        if it % 8 == 0:
            c = c[:len(c) // 2] + g.rs(300) + c[len(c) // 2:]      # a long insertion: wide band
        pairs.append((a, c)); b.add(PC_OP.ALIGN, a, c)
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for (a, c), r, j in zip(pairs, res, jobs):
        s, ops = port.align(a, c)
        assert r[0] == 0 and r[1] == s
        assert var[j["out_off"]:j["out_off"] + r[2]].tobytes() == ops


def test_align_bit_parallel_classes_vs_port(cu, port):
    """compute_alignment through k_align_bp (one job per thread, traceback from stored delta vectors): EST rows around every
    block boundary (63..65, 127..129, 319..321 = the hand-over to the wavefront kernel), empty strings, N / n wildcards and
    lower case on both sides, a byte outside the alphabet (hand-over), a genome piece much longer than the EST (columns
    beyond the thread's share of the pool: hand-over).  Score, length and every alignment column equal the oracle's."""
    g = Gen(31337)
    b, pairs = Batch(), []
    lens = [0, 1, 2, 7, 19, 40, 63, 64, 65, 100, 127, 128, 129, 191, 192, 193, 255, 256, 257, 319, 320, 321, 400]
    for it in range(1500):
        n = lens[it % len(lens)] if it % 3 else g.rnd.randint(0, 90)
        a = g.rs(n)
        alpha = ("ACGT", "ACGTN", "ACGTacgtNn")[it % 3]
        c = g.mutate(a, (0.0, 0.02, 0.08, 0.3)[it % 4], alpha=alpha) if it % 5 else g.rs(g.rnd.randint(0, 120))
        if it % 11 == 0 and len(a) > 3:
            a = a[:2] + b"n" + a[3:]
        if it % 97 == 0 and len(c) > 1:
            c = c[:1] + b"*" + c[2:]
        if it % 131 == 0:
            c = c + g.rs(3000)
        pairs.append((a, c)); b.add(PC_OP.ALIGN, a, c)
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for (a, c), r, j in zip(pairs, res, jobs):
        s, ops = port.align(a, c)
        assert r[0] == 0 and r[1] == s and r[2] == len(ops), (len(a), len(c), r[:3], s, len(ops))
        assert var[j["out_off"]:j["out_off"] + r[2]].tobytes() == ops, (len(a), len(c))


def test_properties_at_scale(cu):
    """Beyond oracle reach (10^4-10^5 nt): score symmetry, ops consistency and identity."""
    g = Gen(5)
    a = g.rs(60000)
    c = g.mutate(a, 0.005, alpha="ACGT")
    s1, ops1 = cu.compute_alignment(a, c)
    s2, ops2 = cu.compute_alignment(c, a)
    assert s1 == s2 == cu.edit_distance(a.replace(b"N", b"A"), c.replace(b"N", b"A")) or s1 == s2
    o = np.frombuffer(ops1, dtype=np.uint8)
    assert (o != 2).sum() == len(a) and (o != 1).sum() == len(c)
    # the cost of the reported path equals the reported score
    i = j = cost = 0
    for x in ops1:
        if x == 0:
            cost += a[i] != c[j] and a[i] not in b"Nn" and c[j] not in b"Nn"; i += 1; j += 1
        elif x == 1:
            cost += 1; i += 1
        else:
            cost += 1; j += 1
    assert cost == s1
    s0, ops0 = cu.compute_alignment(a, a)
    assert s0 == 0 and set(ops0) == {0}
    ln, o1, o2 = cu.find_longest_common_factor_dp(a, a[30000:30040])
    assert ln == 40 and a[o1:o1 + ln] == a[30000:30040] and o2 == 0


def test_gap_packed_classes_vs_port(cu, port):
    """The 16x2-packed gap kernel: every row class (n <= 64 / 128 / 256), the generic fallback (n > 256), ragged pairs
    (odd job counts, very different n and m inside a pair), N wildcards and lower case, against the oracle."""
    g = Gen(777)
    r = g.rnd
    b, cases = Batch(), []
    for it in range(1501):
        n = r.choice([1, 2, 7, 8, 9, 30, 59, 60, 61, 64, 65, 100, 128, 129, 200, 256, 257, 300])
        m = r.choice([1, 2, 3, 50, 100, 199, 200, 201, 260, 400])
        if it % 4 == 0:
            est, gen = g.gap_case()
        else:
            ex1, ex2 = g.rs(n // 2), g.rs(n - n // 2)
            est = g.mutate(ex1 + ex2, r.choice([0, 0.03, 0.15]), alpha="ACGTNn") or b"A"
            gen = (ex1 + g.rs(max(0, m - n)) + ex2)[:max(1, m)] if m >= n else g.rs(m)
            if it % 9 == 0:
                gen = gen.lower()[:len(gen) // 2] + gen[len(gen) // 2:]
        b.add(PC_OP.GAP, est, gen); cases.append((est, gen))
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for (est, gen), rr, j in zip(cases, res, jobs):
        ops, pos = port.gap(est, gen)
        assert rr[0] == 0, (len(est), len(gen), list(rr))
        assert var[j["out_off"]:j["out_off"] + rr[1]].tobytes() == ops, (len(est), len(gen))
        assert list(rr[2:7]) == pos, (len(est), len(gen))


def test_lcs_bit_parallel_vs_port(cu, port):
    """The 64-bit mask form of find_longest_common_factor_dp (s2 <= 64 bytes): N wildcards on both sides, lower case,
    ties between equally long runs (first in (i1, i2) order wins), every s2 length 1..64 and the generic path beyond."""
    g = Gen(31337)
    r = g.rnd
    b, cases = Batch(), []
    for it in range(600):
        l2 = (it % 70) + 1
        l1 = r.choice([1, 2, 5, 31, 32, 33, 63, 64, 65, 255, 256, 257, 300, 1000, 5000])
        s1 = bytearray(g.rs(l1, "ACGT" if it % 3 else "ACGTNn"))
        s2 = bytearray(g.rs(l2, "ACGT" if it % 2 else "ACGTN"))
        if it % 4 == 0 and l1 > l2 + 2:                       # plant s2 (or a piece of it) twice: ties
            piece = s2[: max(1, l2 // 2)]
            for _ in range(2):
                at = r.randint(0, l1 - len(piece))
                s1[at:at + len(piece)] = piece
        if it % 5 == 0:
            s1 = bytearray(bytes(s1).lower()[: len(s1) // 2] + bytes(s1)[len(s1) // 2:])
        if it % 7 == 3:                                       # bytes outside ACGTN / acgtn: the bit-plane form hands the block to the byte form
            tgt = s1 if it % 2 else s2
            tgt[r.randrange(len(tgt))] = r.choice(b"RY*#-x")
        b.add(PC_OP.LCS, bytes(s2), bytes(s1)); cases.append((bytes(s1), bytes(s2)))
    res, _ = cu.run(b)
    for (s1, s2), rr in zip(cases, res):
        assert rr[0] == 0
        assert tuple(rr[1:4]) == port.lcs(s1, s2), (len(s1), len(s2), s1[:80], s2)


def test_bit_parallel_edit_and_kband_vs_port(cu, port):
    """k_myers (one job per thread) and its hand-over list to the wavefront kernel: short and multi-word strings,
    strings above 320 letters, lower-case / N / masked bytes, k = 0, length gaps above k, distances above k inside
    and outside the full-matrix case, equal strings, empty strings — every (ok, edit) and distance equals the port's."""
    import random
    rnd = random.Random(777)
    alphabets = [b"ACGT", b"ACGT", b"ACGTacgtNn", b"ACGTN*#", b"AC"]

    def mutate(s, rate, al):
        out = bytearray()
        for ch in s:
            x = rnd.random()
            if x < rate * 0.5:
                out.append(rnd.choice(al))
            elif x < rate * 0.75:
                out.append(rnd.choice(al)); out.append(ch)
            elif x < rate:
                continue
            else:
                out.append(ch)
        return bytes(out)

    b, chk = Batch(), []
    for it in range(6000):
        al = alphabets[it % len(alphabets)]
        ln = rnd.choice([0, 1, 5, 13, 19, 40, 63, 64, 65, 100, 128, 129, 200, 317, 320, 321, 400])
        a = bytes(rnd.choice(al) for _ in range(ln))
        c = a if it % 11 == 0 else mutate(a, rnd.choice([0.0, 0.02, 0.05, 0.3]), al)
        if it % 13 == 0:
            c = c + bytes(rnd.choice(al) for _ in range(rnd.randint(1, 30)))
        if it % 2:
            a, c = c, a
        b.add(PC_OP.EDIT, a, c); chk.append(("edit", a, c))
        k = rnd.choice([0, 1, 2, 3, 5, 9, 14, 40])
        b.add(PC_OP.KBAND, a, c, p0=k); chk.append(("kband", a, c, k))
    res, _ = cu.run(b)
    for r, c in zip(res, chk):
        assert r[0] == 0, (c, list(r))
        if c[0] == "edit":
            assert r[1] == port.edit(c[1], c[2]), c
        else:
            assert (bool(r[1]), int(r[2])) == port.kband(c[1], c[2], c[3]), c


def test_large_batch_is_ordered_on_the_device(cu, port):
    """>= 65 536 jobs: keys, histogram, scan, scatter and the LCS block prefix run on the GPU (k_order.cu) instead of the
    submitting thread.  Every op, 300 distinct cases cycled to 70 000 jobs; each result equals the port's; a job that
    points outside the arena makes pc_submit fail like the host-side check does."""
    g = Gen(99)
    genome = g.genome(5000)
    cu.genome_upload(genome, 15, 0.2)
    cases = []
    for it in range(300):
        a, c = g.pair(120, it)
        kind = it % 7
        if kind == 0:
            cases.append((PC_OP.EDIT, a, c, {}, ("edit", port.edit(a, c))))
        elif kind == 1:
            k = g.rnd.randint(0, 9)
            cases.append((PC_OP.KBAND, a, c, {"p0": k}, ("kband", port.kband(a, c, k))))
        elif kind == 2:
            est, gen = g.gap_case()
            cases.append((PC_OP.GAP, est, gen, {}, ("gap", port.gap(est, gen))))
        elif kind == 3:
            p, t, me = g.borders_case()
            cases.append((PC_OP.BORDERS, p, t, {"p0": me, "p1": 0, "p2": len(p)}, ("borders", port.borders(p, t, me, 0, len(p)))))
        elif kind == 4:
            cases.append((PC_OP.LCS, c, a, {}, ("lcs", port.lcs(a, c))))
        elif kind == 5:
            cases.append((PC_OP.ALIGN, a, c, {}, ("align", port.align(a, c))))
        else:
            cases.append((PC_OP.AFFIX, a, c, {}, ("affix", port.affix(a, c))))
    b = Batch()
    N = 70000
    for q in range(N):
        op, a, c, kw, _ = cases[q % len(cases)]
        b.add(op, a, c, **kw)
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for q in range(N):
        r, j = res[q], jobs[q]
        kind, want = cases[q % len(cases)][4]
        assert r[0] == 0, (q, kind, list(r))
        if kind == "edit":
            assert r[1] == want
        elif kind == "kband":
            assert (bool(r[1]), int(r[2])) == want
        elif kind == "gap":
            assert var[j["out_off"]:j["out_off"] + r[1]].tobytes() == want[0] and list(r[2:7]) == want[1]
        elif kind == "borders":
            assert bool(r[1]) == want[0] and list(r[2:6]) == want[1]
        elif kind == "lcs":
            assert tuple(r[1:4]) == want
        elif kind == "align":
            assert r[1] == want[0] and var[j["out_off"]:j["out_off"] + r[2]].tobytes() == want[1]
        else:
            assert bool(r[1]) == want[0] and (not want[0] or (r[2], r[3]) == (want[1], want[2]))
    # an out-of-range job in a large batch is refused
    bad = Batch()
    for q in range(N):
        bad.add(PC_OP.EDIT, b"ACGT", b"AGT")
    arena, jb = bad.arrays()
    jb["a_off"][N // 2] = len(arena) + 100
    with pytest.raises(RuntimeError):
        cu.run_arrays(arena, jb, bad.var_bytes)


def test_packed_borders_vs_port(cu, port):
    """k_borders_packed (two matrices per register, rows in registers) and its class boundaries: len_p at 1, 8, 9, 64,
    65, 128, 129, 256 (packed), 257 .. 900 and windows above 1024 columns (row-chunked kernel), windows shorter than p, max_errs 0, single-candidate
    cut ranges, t taken from the device-resident genome, GT..AG planted so that the Burset tie-break decides."""
    import random
    rnd = random.Random(4711)
    g = Gen(5)
    genome = g.genome(8000)
    cu.genome_upload(genome, 15, 0.2)
    b, chk = Batch(), []
    for it in range(2500):
        lp = rnd.choice([1, 2, 7, 8, 9, 20, 40, 63, 64, 65, 100, 128, 129, 200, 256, 257, 400, 511, 513, 900])
        p = g.rs(lp)
        cut = rnd.randint(0, lp)
        mid = g.rs(rnd.choice([0, 0, 5, 60, 300, 1500]))
        if it % 3 == 0 and len(mid) >= 4:
            mid = b"GT" + mid[2:-2] + b"AG"
        t = g.mutate(p[:cut], 0.04) + mid + g.mutate(p[cut:], 0.04)
        if it % 17 == 0:
            t = t[:max(1, lp // 2)]                       # window shorter than p
        if len(t) < 2:
            t += b"AC"
        me = rnd.choice([0, 1, 3, 8, 12, 40, 1500])       # 1500: a window above the packed kernels' 1024 columns when t is long
        lo = rnd.randint(0, lp); hi = rnd.randint(lo, lp)
        if it % 5 == 0:
            lo, hi = 0, lp
        if it % 7 == 0 and len(t) + 2 < len(genome):      # the same t, addressed inside the genome copy on the device
            off = rnd.randint(0, len(genome) - len(t) - 2)
            lt = len(t)
            t = genome[off:off + lt + 2]                  # the reference may read the two bytes after t: here the next genome bytes
            b.add(PC_OP.BORDERS, p, b_in_genome=(off, lt), p0=me, p1=lo, p2=hi)
            chk.append((p, t, lt, me, lo, hi))
        else:
            b.add(PC_OP.BORDERS, p, t, p0=me, p1=lo, p2=hi)
            chk.append((p, t, len(t), me, lo, hi))
    # scores that do not fit 16 bits: handed from the chunked kernel to the wavefront kernel
    p = g.rs(100); t = g.rs(70000)
    b.add(PC_OP.BORDERS, p, t, p0=70000, p1=0, p2=100); chk.append((p, t, len(t), 70000, 0, 100))
    res, _ = cu.run(b)
    import ctypes
    for r, (p, t, lt, me, lo, hi) in zip(res, chk):
        assert r[0] == 0, (p, t, list(r))
        out = (ctypes.c_int * 4)()
        ok = bool(port.lib.po_borders(p, len(p), lo, hi, t, lt, ctypes.c_uint(me), out))
        assert bool(r[1]) == ok and list(r[2:6]) == list(out), (p, t, lt, me, lo, hi, list(r), ok, list(out))


def _meg_shim():
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run(["make", "-s", "-C", os.path.join(root, "tests"), "_build/libmeg_host.so"], check=True)
    L = C.CDLL(os.path.join(root, "tests", "_build", "libmeg_host.so"))
    L.meg_host_record.restype = C.c_longlong
    L.meg_host_record.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    return L


def _meg_cfg(min_intron=60, max_intron=0, max_pairings=80, flags=3, pre=0.6, suf=0.6, freq=0.4):
    import struct
    return struct.pack("<iiIIddd", min_intron, max_intron, max_pairings, flags, pre, suf, freq)


def test_meg_on_device_vs_host_core(cu, port):
    """PC_OP_SEED with PC_SEED_BUILD_MEG: the graph the GPU returns (vertices in list order, ordered adjacency lists, the
    "too complex" flag) equals the record of the same core compiled for the host on the oracle's triples, for spliced ESTs
    over a genome with repeats, every combination of the two simplification switches, several pairing lengths and option
    values; an output region that is too small reports the needed size."""
    L = _meg_shim()
    g = Gen(977)
    genome = g.genome(30000, n_repeats=12)
    cu.genome_upload(genome, 15, 0.2)
    b, chk = Batch(), []
    for it in range(400):
        r = g.rnd
        parts, pos = [], r.randint(0, 4000)
        for _ in range(r.randint(1, 14)):            # spliced EST: exons of 15..160 nt, introns of 3..2500 nt (short ones feed the compaction)
            ln = r.randint(15, 160)
            parts.append(genome[pos:pos + ln])
            pos += ln + r.choice([r.randint(1, 3), r.randint(4, 70), r.randint(70, 2500)])
            if pos >= len(genome) - 200:
                break
        e = g.mutate(b"".join(parts), r.choice([0, 0.01, 0.03]), alpha="ACGT") or b"ACGT"
        cfg = _meg_cfg(min_intron=r.choice([60, 0, 200]), max_intron=r.choice([0, 0, 1500]), max_pairings=r.choice([80, 0, 5]), flags=it % 4,
                       pre=r.choice([0.6, 0.1, 1.0]), suf=r.choice([0.6, 0.1, 1.0]), freq=r.choice([0.4, 0.05]))
        mfl = (15, 15, 16, 19)[it % 4]
        cap = 8 if it % 50 == 7 else 4096
        b.add(PC_OP.SEED, e, b=cfg, p0=mfl, p1=1, out_cap=cap); chk.append((e, cfg, mfl, cap))
    res, var = cu.run(b)
    _, jobs = b.arrays()
    sizes, retries, small = [], 0, 0
    for r, j, (e, cfg, mfl, cap) in zip(res, jobs, chk):
        tri = np.array(port.seed(genome, e, mfl, 0.2), dtype=np.int32).reshape(-1, 3)
        out = np.zeros(1 << 18, dtype=np.int32)
        tri_c = np.ascontiguousarray(tri)
        words = L.meg_host_record(tri_c.ctypes.data, len(tri), len(e), mfl, cfg, out.ctypes.data, len(out))
        assert words > 0
        units = (words + 2) // 3
        if units > cap:
            assert r[0] == -2 and r[1] == units      # PC_E_OUTCAP + the needed 12-byte units
            small += 1
            continue
        assert r[0] == 0 and r[1] == units, (list(r), units)
        got = var[j["out_off"]:j["out_off"] + 4 * words].view(np.int32)
        assert got.tolist() == out[:words].tolist()
        sizes.append(int(out[0])); retries += int(out[2])
    assert max(sizes) >= 12 and small >= 1           # real graphs, and the too-small region happened
