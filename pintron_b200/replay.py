"""The device job stream of a real est-fact run, merged into one batch (bench.py's HBM-resident workload).

libpintron_cuda.so appends every batch it is handed to $PC_CAPTURE (pc_api.cu: u32 njobs, u64 arena_bytes, jobs,
arena).  `capture()` runs the shipped est-fact on a directory with genomic.txt / ests.txt and that variable set;
`merge()` concatenates the recorded batches into one arena + one job array (offsets rebased, output regions laid out
again) and counts the ALGORITHMIC cells of every job with the reference's own formulas, so that a
bench step is exactly "every device job est-fact issues for these ESTs, once".
"""
import os
import subprocess

import numpy as np

from .binding import JOB_DTYPE, PC_B_IN_GENOME, PC_OP

OP_NAMES = ["ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX", "SUFCUT", "PRECUT", "LCS", "SEED"]


def capture(exe, workdir, threads=None, device=None):
    """Run est-fact in `workdir` with PC_CAPTURE set; returns the capture file path."""
    cap = os.path.join(workdir, "jobs.capture")
    env = dict(os.environ, PC_CAPTURE=cap)
    cmd = [exe, "--no-aux-outputs", "--engine", "inproc"]      # the capture hook lives in the library of THIS process
    if threads:
        cmd += ["--threads", str(threads)]
    if device is not None:
        cmd += ["--devices", str(device)]
    p = subprocess.run(cmd, cwd=workdir, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    if p.returncode != 0:
        raise RuntimeError("est-fact (capture run) failed: " + p.stderr.decode("latin1")[-1000:])
    capture.last_log = [l for l in p.stderr.decode("latin1").splitlines() if "scheduler:" in l or "@Timer Total" in l or "thread-seconds" in l]
    return cap


def records(path):
    """Yields (group, jobs, arena) for every record of a capture file: group = running number of the DEVICE batch the
    record was part of (pc_submit_parts writes a marker u32 0xffffffff, u64 nparts before the parts it merged; a plain
    pc_submit record is its own batch)."""
    raw = np.fromfile(path, dtype=np.uint8)
    at, group, left = 0, -1, 0
    while at + 12 <= raw.size:
        n = int(raw[at:at + 4].view("<u4")[0])
        ab = int(raw[at + 4:at + 12].view("<u8")[0])
        at += 12
        if n == 0xffffffff:
            group += 1
            left = ab
            continue
        if left == 0:
            group += 1
        else:
            left -= 1
        j = raw[at:at + n * JOB_DTYPE.itemsize].view(JOB_DTYPE).copy()
        at += n * JOB_DTYPE.itemsize
        yield group, j, raw[at:at + ab]
        at += ab


def _concat(parts, pad):
    """[(jobs, arena)] -> one arena + one job array: offsets rebased, output regions laid out again (4-aligned).  pad =
    spare bytes after every part's arena (the engine leaves 16: BORDERS reads t[len_t], word loads run past the end)."""
    base, arenas, jobs = 0, [], []
    for j, ar in parts:
        j = j.copy()
        j["a_off"] += np.uint32(base)
        in_arena = ((j["flags"] & PC_B_IN_GENOME) == 0) & (j["op"] != PC_OP.SEED)
        j["b_off"][in_arena] += np.uint32(base)
        arenas.append(ar)
        step = len(ar) + pad
        if pad:
            step = (step + 15) & ~15
            arenas.append(np.zeros(step - len(ar), np.uint8))
        base += step
        jobs.append(j)
    jobs = np.concatenate(jobs)
    arena = np.concatenate(arenas) if base else np.zeros(1, np.uint8)
    op = jobs["op"]
    size = np.where((op == PC_OP.ALIGN) | (op == PC_OP.GAP), jobs["out_cap"].astype(np.int64),
                    np.where(op == PC_OP.SEED, 12 * jobs["out_cap"].astype(np.int64), 0))
    size = (size + 3) & ~3
    off = np.concatenate(([0], np.cumsum(size)))
    if off[-1] >= (1 << 32) - 64 or base >= (1 << 32) - 64:
        raise RuntimeError("merged batch passes 4 GiB: capture fewer ESTs")
    jobs["out_off"] = off[:-1].astype(np.uint32)
    return arena, jobs, int(off[-1])


def merge(path, max_arena_bytes=3 << 30):
    """-> (arena uint8[], jobs JOB_DTYPE[], var_bytes, n_records): EVERY recorded job as ONE batch (kernel-level timing at
    full occupancy).  Stops before the merged arena would pass `max_arena_bytes` (job offsets are 32-bit)."""
    parts, tot = [], 0
    for _, j, ar in records(path):
        if tot + len(ar) > max_arena_bytes:
            break
        parts.append((j, ar))
        tot += len(ar)
    if not parts:
        return np.zeros(1, np.uint8), np.zeros(0, JOB_DTYPE), 0, 0
    arena, jobs, var_bytes = _concat(parts, 0)
    return arena, jobs, var_bytes, len(parts)


def device_batches(path):
    """-> list of (arena, jobs, var_bytes, n_lanes): the capture as the DEVICE batches the engine really ran (the lanes it
    merged stay merged, nothing else is), each laid out the way pc_submit_parts lays it out."""
    out, cur, cur_g = [], [], None
    for g, j, ar in records(path):
        if cur and g != cur_g:
            out.append(_concat(cur, 16) + (len(cur),))
            cur = []
        cur_g = g
        if len(j):
            cur.append((j, ar))
    if cur:
        out.append(_concat(cur, 16) + (len(cur),))
    return out


def _equal_pairs(arena, genome, jobs, sel):
    """Boolean per selected job: a == b byte for byte (only jobs with a_len == b_len can be equal)."""
    out = np.zeros(sel.size, dtype=bool)
    j = jobs[sel]
    cand = np.nonzero(j["a_len"] == j["b_len"])[0]
    g = np.frombuffer(genome, dtype=np.uint8)
    CH = 200_000
    for c0 in range(0, cand.size, CH):
        c = cand[c0:c0 + CH]
        ln = j["a_len"][c].astype(np.int64)
        tot = int(ln.sum())
        if tot == 0:
            out[c] = True
            continue
        starts = np.concatenate(([0], np.cumsum(ln)[:-1]))
        within = np.arange(tot, dtype=np.int64) - np.repeat(starts, ln)
        ia = np.repeat(j["a_off"][c].astype(np.int64), ln) + within
        ib = np.repeat(j["b_off"][c].astype(np.int64), ln) + within
        ing = np.repeat((j["flags"][c] & PC_B_IN_GENOME) != 0, ln)
        bb = np.where(ing, g[np.minimum(ib, g.size - 1)], arena[np.minimum(ib, arena.size - 1)])
        eq = arena[ia] == bb
        nz = ln > 0
        res = np.ones(c.size, dtype=bool)
        res[nz] = np.logical_and.reduceat(eq, starts[nz])
        out[c] = res
    return out


def algorithmic_cells(arena, genome, jobs):
    """Cells of the REFERENCE's recurrences per op: ALIGN n*m (0 when the strings are equal),
    KBAND (2k+1)*m when 2k+1 < n else n*m (0 when equal or |n-m| > k), EDIT ls1*ls2, BORDERS 2*t_win*len_p,
    GAP 3*n*m, AFFIX / cuts l1*l2, LCS G_prefix*eplen."""
    op = jobs["op"]
    a, b = jobs["a_len"].astype(np.int64), jobs["b_len"].astype(np.int64)
    cells = {}
    sel = np.nonzero(op == PC_OP.ALIGN)[0]
    if sel.size:
        eq = _equal_pairs(arena, genome, jobs, sel)
        cells["ALIGN"] = int((a[sel] * b[sel])[~eq].sum())
    sel = np.nonzero(op == PC_OP.KBAND)[0]
    if sel.size:
        eq = _equal_pairs(arena, genome, jobs, sel)
        k = jobs["p0"][sel].astype(np.int64)
        n, m = np.maximum(a[sel], b[sel]), np.minimum(a[sel], b[sel])
        c = np.where(2 * k + 1 < n, (2 * k + 1) * m, n * m)
        c[eq | (np.abs(n - m) > k)] = 0
        cells["KBAND"] = int(c.sum())
    for name, code in (("EDIT", PC_OP.EDIT), ("AFFIX", PC_OP.AFFIX), ("SUFCUT", PC_OP.SUFCUT), ("PRECUT", PC_OP.PRECUT), ("LCS", PC_OP.LCS)):
        sel = op == code
        if sel.any():
            cells[name] = int((a[sel] * b[sel]).sum())
    sel = op == PC_OP.BORDERS
    if sel.any():
        t_win = np.minimum(a[sel] + jobs["p0"][sel].astype(np.int64), b[sel])
        cells["BORDERS"] = int((2 * t_win * a[sel]).sum())
    sel = op == PC_OP.GAP
    if sel.any():
        cells["GAP"] = int((3 * a[sel] * b[sel]).sum())
    return cells
