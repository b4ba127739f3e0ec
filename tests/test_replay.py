"""CPU: pintron_b200/replay.py — merging captured pc_submit batches keeps every job's strings and lays the outputs out
again; the algorithmic cell counts follow SURVEY.md §8(d)."""
import numpy as np

from pintron_b200 import replay
from pintron_b200.binding import Batch, PC_OP


def _write_capture(path, batches):
    with open(path, "wb") as f:
        for b in batches:
            arena, jobs = b.arrays()
            f.write(np.uint32(len(jobs)).tobytes())
            f.write(np.uint64(len(b.arena)).tobytes())
            f.write(jobs.tobytes())
            f.write(bytes(b.arena))


def test_merge_rebases_offsets_and_counts_cells(tmp_path):
    genome = b"ACGTACGTTTGACCAGTAGGATCCA" * 40
    b1, b2 = Batch(), Batch()
    b1.add(PC_OP.SEED, b"ACGTTTGACCAGTAGG", p0=15, out_cap=7)
    b1.add(PC_OP.ALIGN, b"ACGTAC", b"ACGTAC")                 # equal strings: 0 cells
    b1.add(PC_OP.ALIGN, b"ACGTAC", b"ACCTACG")                # 6 * 7
    b1.add(PC_OP.KBAND, b"ACGTACGTAA", b"ACGTACGTAT", p0=1)   # 2k+1 = 3 < 10: 3 * 10
    b2.add(PC_OP.GAP, b"ACGTACGT", b"ACGTTTACGT")             # 3 * 8 * 10
    b2.add(PC_OP.BORDERS, b"ACGT", b_in_genome=(5, 100), p0=2, p1=0, p2=4)     # t_win = 6: 2 * 6 * 4
    b2.add(PC_OP.EDIT, b"ACG", b"ACGT")                       # 12
    b2.add(PC_OP.LCS, b"ACGTT", b_in_genome=(0, 200))         # 1000
    b2.add(PC_OP.KBAND, b"ACGT", b"ACGTTTTT", p0=1)           # |n - m| > k: 0
    cap = tmp_path / "jobs.capture"
    _write_capture(cap, [b1, b2])
    arena, jobs, var_bytes, nb = replay.merge(str(cap))
    assert nb == 2 and len(jobs) == 9
    a1, _ = b1.arrays()
    # strings survive the rebasing
    def s_a(j): return bytes(arena[j["a_off"]:j["a_off"] + j["a_len"]])
    def s_b(j): return bytes(arena[j["b_off"]:j["b_off"] + j["b_len"]])
    assert s_a(jobs[0]) == b"ACGTTTGACCAGTAGG" and s_b(jobs[2]) == b"ACCTACG"
    assert s_a(jobs[4]) == b"ACGTACGT" and s_b(jobs[4]) == b"ACGTTTACGT" and jobs[4]["a_off"] >= len(b1.arena)
    assert jobs[5]["b_off"] == 5 and jobs[7]["b_off"] == 0        # genome offsets are not rebased
    # output regions: disjoint, SEED 4-aligned, inside var_bytes
    regs = []
    for j in jobs:
        size = j["out_cap"] if j["op"] in (PC_OP.ALIGN, PC_OP.GAP) else (12 * j["out_cap"] if j["op"] == PC_OP.SEED else 0)
        if size:
            regs.append((int(j["out_off"]), int(j["out_off"]) + int(size)))
            assert j["op"] != PC_OP.SEED or j["out_off"] % 4 == 0
    regs.sort()
    assert all(regs[i][1] <= regs[i + 1][0] for i in range(len(regs) - 1)) and regs[-1][1] <= var_bytes
    cells = replay.algorithmic_cells(arena, genome, jobs)
    assert cells == {"ALIGN": 42, "KBAND": 30, "EDIT": 12, "LCS": 1000, "BORDERS": 48, "GAP": 240}


def test_merge_stops_at_the_arena_limit(tmp_path):
    b = Batch()
    b.add(PC_OP.EDIT, b"A" * 600, b"C" * 400)
    cap = tmp_path / "c"
    _write_capture(cap, [b, b, b])
    _, jobs, _, nb = replay.merge(str(cap), max_arena_bytes=2500)
    assert nb == 2 and len(jobs) == 2
