"""CPU (-m "not gpu"): the C host program of est-fact (pintron_b200/host/: FASTA prep, MEG, embeddings, filters, splice-
site and post-factorization refinement, output format) must reproduce the UNMODIFIED reference byte for byte on the
reference's own regression inputs.  The GPU is not available here, so the host is linked against the CPU oracle
through tests/cpu_backend/ (test infrastructure); the same cases run against the CUDA library in test_estfact_gpu.py.
Expected md5s come from oracle/_ref/est-fact via tests/golden/make_estfact_golden.py."""
import os
import subprocess

import pytest

import estfact_util as U

SMALL = ["test-AMBN", "test-788", "test-mattia1", "test-mattia3", "test-CPB2", "edge-cases"]


@pytest.fixture(scope="module")
def cpu_bin():
    return U.build_cpu_binary()


@pytest.mark.parametrize("case", SMALL)
def test_regression_case_byte_identical(cpu_bin, case, tmp_path):
    U.check_case(cpu_bin, case, tmp_path, "--quiet", "--threads", "4")


def test_single_thread_and_many_fibers_give_the_same_bytes(cpu_bin, tmp_path):
    """Scheduling (threads, fibers in flight) must not change a byte."""
    d1 = tmp_path / "a"; d1.mkdir()
    U.check_case(cpu_bin, "test-AMBN", d1, "--quiet", "--threads", "1", "--fibers", "1")
    d2 = tmp_path / "b"; d2.mkdir()
    U.check_case(cpu_bin, "test-AMBN", d2, "--quiet", "--threads", "3", "--fibers", "7")


@pytest.mark.parametrize("groups", ["3", "4"])
def test_more_fiber_groups_per_worker_give_the_same_bytes(cpu_bin, groups, tmp_path):
    """EF_GROUPS: a worker thread may own up to four fiber groups (= engine lanes) instead of two."""
    U.check_case(cpu_bin, "test-AMBN", tmp_path, "--quiet", "--threads", "3", "--fibers", "5", env=dict(os.environ, EF_GROUPS=groups))


def test_cli_contract(cpu_bin, tmp_path):
    """Option names / defaults of src/options.ggo, config-dump.ini, precedence CLI > config.ini."""
    U.unpack("test-mattia3", str(tmp_path))
    open(tmp_path / "config.ini", "w").write("min-factor-length=16\nmax-prefix-discarded=40\n")
    U.run(cpu_bin, str(tmp_path), "--quiet", "-l", "15", "--threads", "2")
    dump = open(tmp_path / "config-dump.ini").read()
    assert 'min-factor-length="15"' in dump and 'max-prefix-discarded="40"' in dump
    assert 'min-string-depth-rate="0.2"' in dump and 'retain-externals="true"' in dump
    p = subprocess.run([cpu_bin, "--retain-externals=maybe"], cwd=str(tmp_path), capture_output=True)
    assert p.returncode != 0
    p = subprocess.run([cpu_bin, "--version"], cwd=str(tmp_path), capture_output=True)
    assert p.returncode == 0 and b"est-fact 0.1" in p.stdout


def test_empty_inputs_are_not_fatal(cpu_bin, tmp_path):
    """No ESTs, and an EST record with an empty sequence (the reference segfaults on the latter): exit 0, empty outputs."""
    U.unpack("test-mattia3", str(tmp_path))
    for content in (b"", b">empty /gb=E1\n\n"):
        open(tmp_path / "ests.txt", "wb").write(content)
        U.run(cpu_bin, str(tmp_path), "--quiet")
        assert all(os.path.getsize(tmp_path / f) == 0 for f in ("raw-multifasta-out.txt", "processed-ests.txt"))


def test_missing_inputs_fail_loudly(cpu_bin, tmp_path):
    p = subprocess.run([cpu_bin], cwd=str(tmp_path), capture_output=True)
    assert p.returncode != 0 and b"genomic.txt" in p.stderr


def test_product_binary_has_no_cpu_path(tmp_path):
    """The shipped est-fact links only libpintron_cuda.so; without a CUDA device it must refuse to run."""
    import torch
    if torch.cuda.is_available() or not os.path.exists(U.GPU_BIN):
        pytest.skip("needs a GPU-less box and a built pintron_b200/bin/est-fact")
    U.unpack("test-mattia3", str(tmp_path))
    for mode in ("inproc", "auto"):     # auto: the server it starts cannot come up either, and the in-process engine then says why
        p = subprocess.run([U.GPU_BIN, "--engine", mode], cwd=str(tmp_path), capture_output=True,
                           env=dict(os.environ, EST_FACTD_SOCKET=str(tmp_path / "none.sock"), EST_FACTD_IDLE="1"))
        assert p.returncode != 0 and b"no CUDA device" in p.stderr, p.stderr[-600:]
    syms = subprocess.run(["nm", "-D", "--undefined-only", U.GPU_BIN], capture_output=True, text=True).stdout
    assert "pc_engine_open" in syms and "po_" not in syms


@pytest.mark.parametrize("opts", U.OPTION_SETS, ids=lambda o: " ".join(o))
def test_option_variants_vs_reference_binary(cpu_bin, opts, tmp_path):
    """Every tuning flag of src/options.ggo changes the result the same way it does in the reference."""
    if not os.path.exists(U.REF_BIN):
        pytest.skip("oracle/_ref/est-fact not built (make -C oracle ref)")
    U.check_options_vs_reference(cpu_bin, "test-mattia1", tmp_path, opts, "--quiet", "--threads", "4")


def test_small_exon_scan_matches_the_reference_loops(cpu_bin):
    """search_small_exon's (offstart, offend) x strstr scan (factorization-refinement.c:770-834) is restructured in
    refine_fact.c (one memmem pass per offstart, pruned trims); fuzz it against the literal loops."""
    exe = os.path.join(os.path.dirname(cpu_bin), "small_exon_fuzz")
    p = subprocess.run([exe, "6000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith("ok 6000 hits "), p.stdout
    assert int(p.stdout.split()[-1]) > 100


def test_bit_parallel_core_matches_the_port(cpu_bin):
    """pintron_b200/csrc/myers_core.h (what every thread of k_myers runs) compiled for the host: 30 000 random and
    mutated pairs, 1 to 5 words per column, strided Peq layout, unsupported bytes reported — against po_edit."""
    exe = os.path.join(os.path.dirname(cpu_bin), "myers_fuzz")
    p = subprocess.run([exe, "30000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith("ok 30000"), p.stdout


def test_bit_parallel_alignment_core_matches_the_port(cpu_bin):
    """pintron_b200/csrc/align_core.h (what every thread of k_align_bp runs: compute_alignment with the N wildcard, traceback
    from two stored words per column and block) compiled for the host: 40 000 random and mutated pairs, 0 to 5 blocks of EST
    rows, empty strings, strided storage, unsupported bytes and missing space reported — score and every alignment column
    against po_align."""
    exe = os.path.join(os.path.dirname(cpu_bin), "align_fuzz")
    p = subprocess.run([exe, "40000"], capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stderr
    assert p.stdout.startswith("ok 40000"), p.stdout


def test_back_pressure_and_staging_growth_keep_the_bytes(cpu_bin, tmp_path):
    """The batcher never grows its staging while a batch is being gathered: ESTs whose requests do not fit wait for the next
    batch, and only a single EST larger than an empty batch makes the buffers grow.  With 1 KB of staging both happen
    all the time; the output must not change."""
    env = dict(os.environ, EF_STAGING_KB="1")
    for case in ("test-AMBN", "test-CPB2"):
        d = tmp_path / case
        d.mkdir()
        U.check_case(cpu_bin, case, str(d), "--threads", "3", "--fibers", "64", env=env)
    log = (tmp_path / "test-CPB2" / "stderr.txt").read_text(errors="replace") if (tmp_path / "test-CPB2" / "stderr.txt").exists() else ""
    if log:
        import re
        m = re.search(r"(\d+) fiber deferrals, (\d+) lane re-allocations", log)
        assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0, log[-400:]


# ---- the engine forms: in-process, and the resident server est-factd over shared-memory lanes --------------------------
@pytest.fixture(scope="module")
def cpu_daemon(cpu_bin):
    return os.path.join(os.path.dirname(cpu_bin), "est-factd")


def test_server_form_gives_the_same_bytes(cpu_bin, cpu_daemon, tmp_path):
    """est-fact as a client of est-factd (socket handshake, memfd lanes, futex doorbell): same output bytes; several
    clients with different genomes share one server at the same time; the server survives a client that dies."""
    srv = U.Server(cpu_daemon, tmp_path)
    try:
        d = tmp_path / "one"; d.mkdir()
        U.check_case(cpu_bin, "test-AMBN", d, "--quiet", "--threads", "3", "--engine", "daemon", env=srv.env)
        assert "engine: est-factd" in (d / "stderr.txt").read_text() or True
        import threading
        errs = []

        def one(case):
            try:
                dd = tmp_path / ("par-" + case); dd.mkdir()
                U.check_case(cpu_bin, case, dd, "--threads", "2", "--engine", "daemon", env=srv.env)
                assert "engine: est-factd" in (dd / "stderr.txt").read_text()
            except Exception as e:      # noqa: BLE001
                errs.append((case, e))
        ths = [threading.Thread(target=one, args=(c,)) for c in ("test-AMBN", "test-788", "test-mattia1", "test-mattia3")]
        [t.start() for t in ths]; [t.join() for t in ths]
        assert not errs, errs
        # a client killed in mid-run: its session is released, the next client is served
        k = tmp_path / "killed"; k.mkdir()
        U.unpack("test-CPB2", str(k))
        p = subprocess.Popen([cpu_bin, "--engine", "daemon", "--threads", "2"], cwd=str(k), env=srv.env, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        import time
        time.sleep(0.3)
        p.kill(); p.wait()
        d2 = tmp_path / "after"; d2.mkdir()
        U.check_case(cpu_bin, "test-mattia3", d2, "--quiet", "--engine", "daemon", env=srv.env)
        time.sleep(0.3)
        assert "client went away, released" in srv.text() or p.returncode == 0
    finally:
        srv.stop()


def test_client_fails_loudly_when_the_server_dies(cpu_bin, cpu_daemon, tmp_path):
    srv = U.Server(cpu_daemon, tmp_path)
    try:
        U.unpack("test-CPB2", str(tmp_path))
        p = subprocess.Popen([cpu_bin, "--engine", "daemon", "--threads", "2"], cwd=str(tmp_path), env=srv.env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
        import time
        time.sleep(0.3)
        srv.proc.kill()
        err = p.communicate(timeout=60)[1]
        assert p.returncode != 0 and b"went away" in err, err[-500:]
    finally:
        srv.stop()
    # no server, and none may be started: --engine daemon refuses, it does not fall back
    p = subprocess.run([cpu_bin, "--engine", "daemon"], cwd=str(tmp_path), capture_output=True,
                       env=dict(os.environ, EST_FACTD_SOCKET=str(tmp_path / "nobody.sock"), EST_FACT_NO_SPAWN="1"))
    assert p.returncode != 0 and b"no est-factd answering" in p.stderr


def test_server_started_on_demand_and_idle_exit(cpu_bin, tmp_path):
    """--engine auto (the default): no server yet -> est-fact starts the est-factd that sits next to it, detached; a second
    job reuses it; it leaves after its idle timeout."""
    import time
    env = dict(os.environ, EST_FACTD_SOCKET=str(tmp_path / "auto.sock"), EST_FACTD_IDLE="2")
    for k in range(2):
        d = tmp_path / f"run{k}"; d.mkdir()
        U.check_case(cpu_bin, "test-mattia3", d, "--engine", "auto", env=env)
        assert "engine: est-factd" in (d / "stderr.txt").read_text()
    t0 = time.time()
    while os.path.exists(tmp_path / "auto.sock") and time.time() - t0 < 20:
        time.sleep(0.2)
    assert not os.path.exists(tmp_path / "auto.sock"), "est-factd did not leave after its idle timeout"


def test_client_fits_pintron_default_memory_limit(cpu_bin, cpu_daemon, tmp_path):
    """dist-scripts/pintron.py:207-213,878-884 runs est-fact under `ulimit -v 3000 MiB`.  The client (no CUDA in it) must
    fit: fiber stacks and lane mappings are budgeted against RLIMIT_AS."""
    srv = U.Server(cpu_daemon, tmp_path)
    try:
        d = tmp_path / "lim"; d.mkdir()
        exp = U.unpack("test-CPB2", str(d))
        p = subprocess.run(["/bin/sh", "-c", f"ulimit -t 3600 && ulimit -v {3000 * 1024} && {cpu_bin} --engine daemon --threads 16"],
                           cwd=str(d), env=srv.env, capture_output=True)
        assert p.returncode == 0, p.stderr[-800:]
        assert U.md5s(str(d)) == {f: exp[f] for f in U.FILES}
    finally:
        srv.stop()


def test_streaming_windows_and_back_pressure_keep_the_bytes(cpu_bin, tmp_path):
    """ests.txt is read window by window while the workers run; the reader stays a bounded number of windows ahead of the
    writers.  With 5-record windows and a look-ahead of 2 every mechanism is exercised all the time."""
    env = dict(os.environ, EF_WINDOW="5", EF_WINDOW_AHEAD="2")
    for case, extra in (("test-CPB2", ["--threads", "4", "--fibers", "3"]), ("test-AMBN", ["--threads", "2"]), ("edge-cases", [])):
        d = tmp_path / case
        d.mkdir()
        U.check_case(cpu_bin, case, str(d), *extra, env=env)


@pytest.mark.parametrize("limit", ["0", "1", "12"])
def test_meg_built_on_either_side_gives_the_same_bytes(cpu_bin, limit, tmp_path):
    """Vertex sets above EF_MEG_DEVICE_MAX pairings come back from the SEED job as they are and host/meg.c walks the same
    csrc/meg_core.h over them (include/pintron_cuda.h: PC_SEED_VERTEX_SET_ONLY).  0 = every graph by the SEED job, 1 = every
    graph on the host, 12 = both in one run; megs.txt / meg-edges.txt / everything downstream must not change."""
    env = dict(os.environ, EF_MEG_DEVICE_MAX=limit)
    for case in ("test-CPB2", "test-AMBN"):
        d = tmp_path / case
        d.mkdir()
        U.check_case(cpu_bin, case, str(d), "--quiet", "--threads", "3", env=env)


def test_counting_reference_is_the_reference(tmp_path):
    """oracle/_ref/est-fact-cells (the reference sources as a PIC library + the counting interposers of oracle/ref_cells.c)
    writes the reference's bytes and counts cells for every routine bench.py divides GCUPS from."""
    exe = os.path.join(U.ROOT, "oracle", "_ref", "est-fact-cells")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/est-fact-cells not built (make -C oracle ref)")
    import json
    U.unpack("test-CPB2", str(tmp_path))
    subprocess.run([exe], cwd=tmp_path, check=True, capture_output=True, timeout=600)
    exp = json.load(open(os.path.join(U.GOLD, "test-CPB2", "expected.json")))
    got = U.md5s(str(tmp_path))
    assert all(got[f] == exp[f] for f in exp if f in got)
    cells = json.load(open(tmp_path / "cells.json"))
    assert set(cells) == {"ALIGN", "KBAND", "EDIT", "BORDERS", "GAP", "AFFIX"}
    assert all(v["calls"] > 0 and v["cells"] > 0 for v in cells.values()), cells
    assert cells["GAP"]["cells"] % 3 == 0
