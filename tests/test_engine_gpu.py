"""GPU (-m gpu): the batch engine (include/pintron_engine.h) — merged multi-part batches, lanes, the resident server
est-factd — against the single-batch path and the reference outputs.  Merging must never change a result."""
import os
import shutil
import subprocess
import threading

import numpy as np
import pytest

import estfact_util as U
import pintron_b200
from pintron_b200 import Batch, PC_OP
from pintron_b200 import binding
from oracle.binding import Port
from util_cases import Gen

pytestmark = pytest.mark.gpu
DAEMON = os.path.join(U.ROOT, "pintron_b200", "bin", "est-factd")


@pytest.fixture(scope="module")
def cu():
    c = pintron_b200.Cuda(0)
    yield c
    c.close()


def _mixed_batches(g, genome, nparts, per_part):
    bs = []
    for q in range(nparts):
        b = Batch()
        for it in range(per_part):
            a, c = g.pair(180, it + q)
            b.add(PC_OP.ALIGN, a, c)
            b.add(PC_OP.EDIT, a, c)
            b.add(PC_OP.KBAND, a, c, p0=g.rnd.randint(0, 12))
            b.add(PC_OP.AFFIX, a, c)
            b.add(PC_OP.LCS, c[:40] or b"A", a)
            p, t, me = g.borders_case()
            b.add(PC_OP.BORDERS, p, t, p0=me, p1=0, p2=len(p))
            est, gen = g.gap_case()
            b.add(PC_OP.GAP, est, gen)
            b.add(PC_OP.SEED, g.est_from(genome, it), p0=15, out_cap=1024)
            if it % 5 == 0:      # genome-side strings by reference
                off = g.rnd.randint(0, len(genome) - 300)
                b.add(PC_OP.ALIGN, g.mutate(genome[off:off + 200], 0.03) or b"A", b_in_genome=(off, 200))
        bs.append(b)
    return bs


def _same(r1, v1, r2, v2, batch):
    _, jobs = batch.arrays()
    assert (r1 == r2).all(), np.nonzero((r1 != r2).any(axis=1))[0][:5]
    for r, j in zip(r1, jobs):
        if j["op"] in (PC_OP.ALIGN, PC_OP.GAP):
            n = r[2] if j["op"] == PC_OP.ALIGN else r[1]
            assert (v1[j["out_off"]:j["out_off"] + n] == v2[j["out_off"]:j["out_off"] + n]).all()
        elif j["op"] == PC_OP.SEED and r[0] == 0:
            assert (v1[j["out_off"]:j["out_off"] + 12 * r[1]] == v2[j["out_off"]:j["out_off"] + 12 * r[1]]).all()


def test_parts_equal_single_batches(cu):
    """pc_submit_parts: 7 ragged parts (one empty) as one device batch == each part submitted alone."""
    g = Gen(9001)
    genome = g.genome(8000)
    cu.genome_upload(genome, 15, 0.2)
    bs = _mixed_batches(g, genome, 6, 25)
    bs.insert(3, Batch())
    bs.append(_mixed_batches(g, genome, 1, 3)[0])
    merged = cu.run_parts(bs)
    for b, (res, var) in zip(bs, merged):
        if not b.jobs:
            continue
        r1, v1 = cu.run(b)
        assert (r1[:, 0] == 0).all()
        _same(r1, v1, res, var, b)


def test_engine_lanes_merge_and_match(cu):
    """Lanes posted to an in-process engine (the host's path) give the single-batch results; lanes posted together are
    merged into fewer device batches than lanes."""
    g = Gen(9002)
    genome = g.genome(8000)
    cu.genome_upload(genome, 15, 0.2)
    bs = _mixed_batches(g, genome, 8, 20)
    eng = binding.Engine((0,), 64 << 20)
    try:
        ses = eng.open(genome, 8, 1 << 20, 8192, 1 << 20)
        for rnd in range(3):
            for k, b in enumerate(bs):
                ses.post(k, b)
            for k, b in enumerate(bs):
                res, var = ses.wait(k, len(b.jobs), b.var_bytes)
                r1, v1 = cu.run(b)
                _same(r1, v1, res, var, b)
        st = ses.close()
        assert st.lanes_merged == 24 and st.batches <= 24 and st.jobs == 3 * sum(len(b.jobs) for b in bs)
        assert st.launches > 0
    finally:
        eng.close()


def test_single_job_larger_than_a_quarter_of_the_pool(cu):
    """a job whose scratch exceeds pool/4 (64 MB) used to loop through 64 retry rounds and come back as
    PC_E_POOL.  A generic GAP job with (n+1)(m+1) > 64 MB, an ALIGN of two unrelated 6 kb strings (full-width band):
    retried with ONE warp owning the whole pool, grown when even that is too small."""
    port = Port()
    g = Gen(9003)
    b = Batch()
    est, gen = g.rs(9000), g.rs(9000)                   # 81 MB of direction bytes
    b.add(PC_OP.GAP, est, gen)
    a, c = g.rs(300), g.rs(280)
    b.add(PC_OP.EDIT, a, c)
    res, var = cu.run(b)
    assert (res[:, 0] == 0).all(), res[:, 0]
    ops, pos = port.gap(est, gen)
    _, jobs = b.arrays()
    assert var[jobs[0]["out_off"]:jobs[0]["out_off"] + res[0, 1]].tobytes() == ops and list(res[0, 2:7]) == pos
    assert res[1, 1] == port.edit(a, c)


def test_gap_job_beyond_the_packed_kernel_columns(cu):
    """n <= 64 with m in (3199, 4096] used to be classed for the packed kernel and fail the launch (shared
    memory); now routed to the generic kernel.  Also m just inside the packed bound."""
    port = Port()
    g = Gen(9004)
    b, cases = Batch(), []
    for m in (2990, 3000, 3001, 3500, 4096, 4500):
        est = g.rs(50)
        gen = g.rs(m // 2) + g.mutate(est, 0.05) + g.rs(m - m // 2)
        b.add(PC_OP.GAP, est, gen); cases.append((est, gen))
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for (est, gen), r, j in zip(cases, res, jobs):
        assert r[0] == 0, (len(gen), list(r))
        ops, pos = port.gap(est, gen)
        assert var[j["out_off"]:j["out_off"] + r[1]].tobytes() == ops and list(r[2:7]) == pos, len(gen)


# ---- est-fact through the server --------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def server(tmp_path_factory):
    assert os.path.exists(DAEMON), "pintron_b200/bin/est-factd is not built"
    srv = U.Server(DAEMON, tmp_path_factory.mktemp("efd"))
    yield srv
    srv.stop()


@pytest.mark.parametrize("case", ["test-AMBN", "test-CPB2", "test_gtf7", "edge-cases"])
def test_server_form_byte_identical(server, case, tmp_path):
    U.check_case(U.GPU_BIN, case, tmp_path, "--engine", "daemon", env=server.env)
    assert "engine: est-factd" in (tmp_path / "stderr.txt").read_text()


def test_server_concurrent_clients_and_staging_growth(server, tmp_path):
    """Four est-fact clients with four genomes on one server at once (their lanes share the GPU, batches stay per genome);
    one of them with 1 KB lanes, so that lane re-allocation over the socket happens all the time."""
    errs = []

    def one(case, env):
        try:
            d = tmp_path / case; d.mkdir()
            U.check_case(U.GPU_BIN, case, d, "--threads", "3", "--engine", "daemon", env=env)
        except Exception as e:      # noqa: BLE001
            errs.append((case, repr(e)[:600]))
    jobs = [("test-AMBN", server.env), ("test-788", server.env), ("test-mattia1", server.env), ("test-CPB2", dict(server.env, EF_STAGING_KB="1"))]
    ths = [threading.Thread(target=one, args=j) for j in jobs]
    [t.start() for t in ths]; [t.join() for t in ths]
    assert not errs, errs
    assert "lane re-allocations" in (tmp_path / "test-CPB2" / "stderr.txt").read_text()


def test_client_under_pintron_default_ulimit(server, tmp_path):
    """dist-scripts/pintron.py:878-884 starts est-fact as `ulimit -t T && ulimit -v V && est-fact` with V = 3000 MiB by
    default — too little for a CUDA context, enough for the CUDA-free client of est-factd."""
    exp = U.unpack("test-CPB2", str(tmp_path))
    p = subprocess.run(["/bin/sh", "-c", f"ulimit -t 3600 && ulimit -v {3000 * 1024} && {U.GPU_BIN} --engine daemon"],
                       cwd=str(tmp_path), env=server.env, capture_output=True)
    assert p.returncode == 0, p.stderr[-800:]
    assert U.md5s(str(tmp_path)) == {f: exp[f] for f in U.FILES}


def test_product_client_refuses_the_cpu_stand_in_server(tmp_path):
    """The test suite's CPU stand-in server must never serve the shipped est-fact: the handshake names the backend."""
    fake = os.path.join(U.HERE, "_build", "est-factd")
    if not os.path.exists(fake):
        subprocess.run(["make", "-s", "-C", U.HERE], check=True)
    srv = U.Server(fake, tmp_path)
    try:
        U.unpack("test-mattia3", str(tmp_path))
        p = subprocess.run([U.GPU_BIN, "--engine", "daemon"], cwd=str(tmp_path), env=srv.env, capture_output=True)
        assert p.returncode != 0 and b"oracle-test" in p.stderr, p.stderr[-500:]
    finally:
        srv.stop()


def test_in_process_engine_two_devices(tmp_path):
    """--devices 0,1 in one process (threads dealt round-robin to the GPUs; per-device kernel attributes) with a
    window above 48 KB of dynamic shared memory (--suff-pref-length-intron 800 -> GAP m > 800)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    if not os.path.exists(U.REF_BIN):
        pytest.skip("oracle/_ref/est-fact not built")
    U.check_options_vs_reference(U.GPU_BIN, "test-CPB2", tmp_path, ["--suff-pref-length-intron", "800"], "--quiet", "--devices", "0,1", "--threads", "4")


def test_large_intron_window_option_vs_reference(tmp_path):
    """--suff-pref-length-intron 2000: GAP jobs with m ~ 4060 columns (beyond the packed kernel) — VERDICT r1 #9."""
    if not os.path.exists(U.REF_BIN):
        pytest.skip("oracle/_ref/est-fact not built")
    U.check_options_vs_reference(U.GPU_BIN, "test-mattia1", tmp_path, ["--suff-pref-length-intron", "2000"], "--quiet")


def test_guard_bands_stay_intact(tmp_path):
    """PC_GUARD=1: every device buffer exactly sized between 0xA5 guard bands, buffers pre-filled with 0xA5, bands checked
    after every batch (compute-sanitizer is closed on this GPU pool; this is the bounds check of our own).  The C-ABI
    parity tests and the program on three regression cases must pass unchanged: no kernel writes out of bounds, no
    result depends on bytes nobody wrote."""
    import sys
    env = dict(os.environ, PC_GUARD="1")
    p = subprocess.run([sys.executable, "-m", "pytest", os.path.join(U.HERE, "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-p", "no:cacheprovider",
                        "-k", "mixed_batch or golden or packed or bit_parallel or lcs or edge_cases or long_alignment"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert p.returncode == 0, p.stdout[-2500:]
    for case in ("test-AMBN", "test-CPB2", "test_gtf7"):
        d = tmp_path / case; d.mkdir()
        U.check_case(U.GPU_BIN, case, d, "--quiet", "--threads", "6", "--engine", "inproc", env=env)


def test_kband_ok_only_flag_and_seed_size_classes(cu):
    """PC_KBAND_OK_ONLY (what the est-fact host sets: the reference's call sites read the boolean only): same `ok` as the exact
    form on every job, and identical (ok, edit) whenever ok.  SEED jobs of 300 nt to 7 kbp in one batch (three size classes,
    scratch slots sized by the longest read of a class) == the port's vertex sets."""
    import random
    port = Port()
    rnd = random.Random(4711)
    g = Gen(4711)
    b1, b2, cases = Batch(), Batch(), []
    for it in range(3000):
        ln = rnd.choice([20, 60, 64, 65, 100, 130, 200, 300, 320, 400])
        a = g.rs(ln)
        c = g.mutate(a, rnd.choice([0.0, 0.02, 0.05, 0.1, 0.3])) or b"A"
        k = rnd.randint(0, 12)
        b1.add(PC_OP.KBAND, a, c, p0=k)
        b2.add(PC_OP.KBAND, a, c, p0=k)
        cases.append((a, c, k))
    b2.jobs = [(op, flags | 4, *rest) for (op, flags, *rest) in b2.jobs]
    r1, _ = cu.run(b1)
    r2, _ = cu.run(b2)
    assert (r1[:, 0] == 0).all() and (r2[:, 0] == 0).all()
    assert (r1[:, 1] == r2[:, 1]).all()
    okm = r1[:, 1] != 0
    assert (r1[okm, 2] == r2[okm, 2]).all()
    for (a, c, k), r in list(zip(cases, r2))[:400]:
        assert bool(r[1]) == port.kband(a, c, k)[0]
    genome = g.genome(60000)
    cu.genome_upload(genome, 15, 0.2)
    b = Batch()
    ests = []
    for ln in (300, 800, 1024, 1025, 2000, 3072, 3073, 5000, 7000, 400, 6000):
        off = rnd.randint(0, len(genome) - ln - 1)
        e = g.mutate(genome[off:off + ln], 0.01)
        ests.append(e)
        b.add(PC_OP.SEED, e, p0=15, out_cap=8192)
    res, var = cu.run(b)
    _, jobs = b.arrays()
    for e, r, j in zip(ests, res, jobs):
        assert r[0] == 0, (len(e), list(r))
        tri = var[j["out_off"]:j["out_off"] + 12 * r[1]].view(np.int32).reshape(-1, 3)
        assert [tuple(map(int, x)) for x in tri] == port.seed(genome, e, 15, 0.2), len(e)
