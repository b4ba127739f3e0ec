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
    p = subprocess.run([U.GPU_BIN], cwd=str(tmp_path), capture_output=True)
    assert p.returncode != 0 and b"no CUDA device" in p.stderr
    syms = subprocess.run(["nm", "-D", "--undefined-only", U.GPU_BIN], capture_output=True, text=True).stdout
    assert "pc_submit" in syms and "po_" not in syms


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
        m = re.search(r"(\d+) fiber deferrals, (\d+) staging re-allocations", log)
        assert m and int(m.group(1)) > 0 and int(m.group(2)) > 0, log[-400:]
