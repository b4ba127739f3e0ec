"""GPU (-m gpu): the shipped est-fact (C host + libpintron_cuda.so through the C ABI) against the UNMODIFIED reference,
byte for byte: (1) every committed regression fixture (md5s written by oracle/_ref/est-fact,
tests/golden/make_estfact_golden.py); (2) synthetic inputs of the bench shape, with the reference binary run beside it
on the box's host cores; (3) scheduling independence (threads / fibers) and EST sharding (shards concatenate to the
single run: the multi-GPU contract)."""
import os
import shutil
import subprocess

import pytest

import estfact_util as U
from pintron_b200.synth import Synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_bin():
    assert os.path.exists(U.GPU_BIN), "pintron_b200/bin/est-fact is not built (python __graft_entry__.py)"
    return U.GPU_BIN


@pytest.mark.parametrize("case", U.cases())
def test_regression_case_byte_identical(gpu_bin, case, tmp_path):
    U.check_case(gpu_bin, case, tmp_path, "--quiet")


def _write_synth(d, name, reads, seed=None):
    s = Synth(name, reads=reads, seed=seed)
    open(os.path.join(d, "genomic.txt"), "wb").write(s.genome_fasta())
    open(os.path.join(d, "ests.txt"), "wb").write(s.ests_fasta(0, reads))


@pytest.mark.parametrize("name,reads", [("tiny", 64), ("C3", 1500), ("C4mini", 400), ("C5mini", 16)])
def test_synthetic_vs_reference_binary(gpu_bin, name, reads, tmp_path):
    if not os.path.exists(U.REF_BIN):
        pytest.skip("oracle/_ref/est-fact not built")
    a, b = tmp_path / "ref", tmp_path / "ours"
    a.mkdir(); b.mkdir()
    _write_synth(str(a), name, reads)
    for f in ("genomic.txt", "ests.txt"):
        shutil.copy(a / f, b / f)
    U.run(U.REF_BIN, str(a))
    U.run(gpu_bin, str(b), "--quiet")
    assert U.md5s(str(a)) == U.md5s(str(b))


def test_scheduling_and_sharding_do_not_change_bytes(gpu_bin, tmp_path):
    whole = tmp_path / "whole"; whole.mkdir()
    exp = U.unpack("test-CPB2", str(whole))
    U.run(gpu_bin, str(whole), "--quiet", "--threads", "1", "--fibers", "3")
    assert U.md5s(str(whole)) == {f: exp[f] for f in U.FILES}
    # two shards of ests.txt, outputs concatenated
    recs = open(whole / "ests.txt", "rb").read().split(b">")[1:]
    half = len(recs) // 2
    cat = {f: b"" for f in U.FILES}
    for k, part in enumerate((recs[:half], recs[half:])):
        d = tmp_path / f"shard{k}"; d.mkdir()
        shutil.copy(whole / "genomic.txt", d / "genomic.txt")
        open(d / "ests.txt", "wb").write(b"".join(b">" + r for r in part))
        U.run(gpu_bin, str(d), "--quiet", "--threads", "5")
        for f in U.FILES:
            cat[f] += open(d / f, "rb").read()
    for f in U.FILES:
        assert cat[f] == open(whole / f, "rb").read(), f


@pytest.mark.parametrize("limit", ["0", "1", "12"])
def test_meg_built_on_either_side_gives_the_same_bytes(gpu_bin, limit, tmp_path):
    """EF_MEG_DEVICE_MAX: 0 = every graph by the SEED kernel's MEG stage, 1 = every graph on the host from the vertex set the
    kernel hands back (PC_SEED_VERTEX_SET_ONLY), 12 = both in one run.  Same bytes as the reference either way."""
    env = dict(os.environ, EF_MEG_DEVICE_MAX=limit)
    for case in ("test-CPB2", "test_gtf7"):
        d = tmp_path / case
        d.mkdir()
        U.check_case(gpu_bin, case, str(d), "--quiet", env=env)


@pytest.mark.parametrize("opts", U.OPTION_SETS, ids=lambda o: " ".join(o))
def test_option_variants_vs_reference_binary(gpu_bin, opts, tmp_path):
    if not os.path.exists(U.REF_BIN):
        pytest.skip("oracle/_ref/est-fact not built")
    U.check_options_vs_reference(gpu_bin, "test-CPB2", tmp_path, opts, "--quiet")
