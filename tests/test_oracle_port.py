"""CPU: the oracle port (oracle/port/dp_port.c) against the golden vectors produced by the UNMODIFIED
reference (tests/golden/make_dp_golden.py), and — where oracle/_ref was built — against the reference itself
on fresh seeded inputs."""
import pytest

from oracle.binding import Port, Ref, ops_to_rows
from util_cases import Gen, enc, golden


@pytest.fixture(scope="module")
def port():
    return Port()


@pytest.fixture(scope="module")
def gold():
    return golden()


def test_align_golden(port, gold):
    for c in gold["align"]:
        est, gen = enc(c["est"]), enc(c["gen"])
        score, ops = port.align(est, gen)
        assert score == c["score"]
        assert ops_to_rows(ops, est, gen) == (enc(c["est_row"]), enc(c["gen_row"]))


def test_edit_kband_golden(port, gold):
    for c in gold["edit"]:
        assert port.edit(enc(c["a"]), enc(c["b"])) == c["dist"]
    for c in gold["kband"]:
        assert port.kband(enc(c["a"]), enc(c["b"]), c["k"]) == (c["ok"], c["edit"])


def test_borders_golden(port, gold):
    for c in gold["borders"]:
        assert port.borders(enc(c["p"]), enc(c["t"]), c["max_errs"]) == (c["ok"], c["out"])


def test_gap_golden(port, gold):
    for c in gold["gap"]:
        est, gen = enc(c["est"]), enc(c["gen"])
        ops, pos = port.gap(est, gen)
        assert ops_to_rows(ops, est, gen) == (enc(c["est_row"]), enc(c["gen_row"]))
        assert pos == c["pos"]


def test_affix_cuts_lcs_golden(port, gold):
    for c in gold["affix"]:
        assert list(port.affix(enc(c["est"]), enc(c["gen"]))) == c["out"]
    for c in gold["suffix_cut"]:
        assert list(port.suffix_cut(enc(c["a"]), enc(c["b"]))) == c["out"]
    for c in gold["prefix_cut"]:
        assert list(port.prefix_cut(enc(c["a"]), enc(c["b"]))) == c["out"]
    for c in gold["lcs"]:
        assert list(port.lcs(enc(c["s1"]), enc(c["s2"]))) == c["out"]


def test_burset_golden(port, gold):
    # reference unit tests pin the same table: test/refine-intron_test.c:148-922
    for c in gold["burset"]:
        assert port.burset(enc(c["donor"]), enc(c["acceptor"])) == c["freq"]
    assert port.burset(b"GT", b"AG") == 200 and port.burset(b"GC", b"AG") == 126 and port.burset(b"gt", b"ag") == 200


def test_seed_golden(port, gold):
    s = gold["seed"]
    g = enc(s["genome"])
    for c in s["cases"]:
        assert port.seed(g, enc(c["est"]), c["mfl"], s["rate"]) == [tuple(x) for x in c["pairings"]]


def test_edge_cases(port):
    assert port.align(b"", b"ACG") == (3, b"\x02\x02\x02")
    assert port.align(b"ACG", b"") == (3, b"\x01\x01\x01")
    assert port.align(b"ANG", b"ACG")[0] == 0
    assert port.edit(b"", b"") == 0 and port.edit(b"ANG", b"ACG") == 1
    assert port.kband(b"ACGT", b"ACGT", 0) == (True, 0)
    assert port.kband(b"ACGT", b"ACGA", 0) == (False, 1)
    assert port.kband(b"ACGTACGTAC", b"ACG", 2) == (False, 7)
    assert port.lcs(b"", b"ACG") == (0, 0, 0)
    assert port.lcs(b"TTACGTT", b"GACGA") == (3, 2, 1)
    assert port.lcs(b"TTANGTT", b"GACGA") == (3, 2, 1)


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_port_matches_reference_fuzz(port):
    ref = Ref()
    g = Gen(77)
    for it in range(400):
        a, b = g.pair(140, it)
        s, ops = port.align(a, b)
        rs, ra, rb = ref.align(a, b)
        assert s == rs and ops_to_rows(ops, a, b) == (ra, rb)
        assert port.edit(a, b) == ref.edit(a, b) == ref.compute_edit(a, b)
        k = g.rnd.randint(0, 12)
        assert port.kband(a, b, k) == ref.kband(a, b, k)
        assert port.suffix_cut(a, b) == ref.suffix_cut(a, b)
        assert port.prefix_cut(a, b) == ref.prefix_cut(a, b)
        assert port.affix(a, b) == ref.affix(a, b)
        assert port.lcs(a, b) == ref.lcs(a, b)
        p, t, me = g.borders_case()
        assert port.borders(p, t, me) == ref.borders(p, t, me)
        est, gen = g.gap_case()
        ops, pos = port.gap(est, gen)
        ra, rb, rpos = ref.gap(est, gen)
        assert ops_to_rows(ops, est, gen) == (ra, rb) and pos == rpos


@pytest.mark.skipif(not Ref.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_seed_matches_reference_fuzz(port):
    ref = Ref()
    g = Gen(78)
    genome = g.genome(4000)
    ix = ref.index(genome, 15, 0.2)
    for it in range(40):
        est = g.est_from(genome, it)
        mfl = (15, 16, 19)[it % 3]
        assert port.seed(genome, est, mfl, 0.2) == ref.seed(ix, est, mfl)
