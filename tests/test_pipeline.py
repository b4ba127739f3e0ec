"""The drop-in behind the reference's own driver: dist-scripts/pintron.py, UNMODIFIED and
with its DEFAULT limits (`ulimit -t 3600 && ulimit -v 3000 MiB`, pintron.py:207-213,878-884), finds `est-fact` in --bin-dir,
runs it, and feeds its raw-multifasta-out.txt / processed-ests.txt to the reference's own min-factorization
(src/main-min-factorization.c:46, src/io-factorizations.c:194-241), intron-agreement, compact-compositions,
maximal-transcripts and cds-annotation.  With our est-fact in that directory the pipeline must produce the same
full.json as with the reference's est-fact, and the 13 introns of regressionTest/test-AMBN/referenceOutput/full.json.

The server est-factd runs OUTSIDE the job's ulimit (a CUDA context does not fit 3000 MiB of address space); est-fact
itself is a CUDA-free client and must fit.  CPU form (-m "not gpu"): the host code over the CPU stand-in engine; GPU
form: the shipped binaries."""
import json
import os
import shutil
import subprocess

import pytest

import estfact_util as U

PIPE = os.path.join(U.ROOT, "oracle", "_ref", "pipeline-bin")
STAGES = ["min-factorization", "intron-agreement", "compact-compositions", "maximal-transcripts", "cds-annotation"]
# regressionTest/test-AMBN/referenceOutput/full.json (older key names: "relative start", "relative end", "number supporting EST")
AMBN_GOLDEN_INTRONS = None


def _golden_introns():
    path = os.path.join(U.GOLD, "test-AMBN", "golden_introns.json")
    return [tuple(x) for x in json.load(open(path))]


def _bin_dir(tmp, est_fact):
    d = os.path.join(str(tmp), "bin")
    os.makedirs(d)
    os.symlink(est_fact, os.path.join(d, "est-fact"))
    for s in STAGES:
        os.symlink(os.path.join(PIPE, s), os.path.join(d, s))
    return d


def _run_pintron(bindir, work, env):
    os.makedirs(work)
    U.unpack("test-AMBN", work)
    p = subprocess.run(["python3", os.path.join(PIPE, "pintron"), "-k", "--bin-dir=" + bindir, "-g", "genomic.txt", "-s", "ests.txt",
                        "--output=full.json", "--gtf=out.gtf", "--organism=human", "--gene=AMBN"],
                       cwd=work, env=env, capture_output=True, timeout=900)
    log = ""
    for f in ("pintron-pipeline-log.txt", "pintron-log.txt"):
        if os.path.exists(os.path.join(work, f)):
            log += open(os.path.join(work, f), errors="replace").read()[-3000:]
    assert p.returncode == 0, (p.stderr.decode("latin1")[-1500:], log[-2500:])
    return json.load(open(os.path.join(work, "full.json"))), log


def _introns(full):
    return sorted((v["relative_start"], v["relative_end"], v["number_of_supporting_transcripts"]) for v in full["introns"].values())


def _check(tmp_path, est_fact, daemon):
    if not os.path.exists(os.path.join(PIPE, "pintron")) or not os.path.exists(U.REF_BIN):
        pytest.skip("oracle/_ref/pipeline-bin not built (make -C oracle ref)")
    ref_full, _ = _run_pintron(_bin_dir(tmp_path / "r", U.REF_BIN), str(tmp_path / "r" / "w"), dict(os.environ))
    srv = U.Server(daemon, tmp_path)                       # started here = outside pintron.py's ulimit, as deployed
    try:
        ours_full, log = _run_pintron(_bin_dir(tmp_path / "o", est_fact), str(tmp_path / "o" / "w"), srv.env)
    finally:
        srv.stop()
    assert "session 1 opened" in srv.text(), "est-fact did not go through est-factd"
    for f in ("raw-multifasta-out.txt", "processed-ests.txt", "out-agree.txt", "out-after-intron-agree.txt", "predicted-introns.txt"):
        assert open(tmp_path / "o" / "w" / f, "rb").read() == open(tmp_path / "r" / "w" / f, "rb").read(), f
    assert ours_full["introns"] == ref_full["introns"] and ours_full["isoforms"] == ref_full["isoforms"]
    assert _introns(ours_full) == sorted(_golden_introns())


def test_pintron_py_with_default_limits_cpu(tmp_path):
    cpu_bin = U.build_cpu_binary()
    _check(tmp_path, cpu_bin, os.path.join(os.path.dirname(cpu_bin), "est-factd"))


@pytest.mark.gpu
def test_pintron_py_with_default_limits_gpu(tmp_path):
    _check(tmp_path, U.GPU_BIN, os.path.join(U.ROOT, "pintron_b200", "bin", "est-factd"))
