"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/pintron_cuda.h declares.
No compute calls (no GPU here); with no device the library must fail loudly, not fall back."""
import ctypes
import os
import re
import subprocess

import pytest

import pintron_b200
from pintron_b200 import binding

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(binding.library_path()):
        binding.build_library()
    return binding.load_library()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pintron_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pc_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pintron_cuda.h but not exported"
    assert set(binding.EXPORTS) <= set(syms)


def test_job_struct_layout():
    assert ctypes.sizeof(binding.pc_job) == 44 == binding.JOB_DTYPE.itemsize


def test_library_is_sm100a():
    out = subprocess.run(["cuobjdump", "-lelf", binding.library_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        pintron_b200.Cuda()
    assert lib.pc_ctx_create(0) is None
    assert lib.pc_last_error()


def test_batch_packing():
    b = pintron_b200.Batch()
    i = b.add(pintron_b200.PC_OP.ALIGN, b"ACGT", b"ACG")
    j = b.add(pintron_b200.PC_OP.SEED, b"ACGTACGTACGTACGTA", p0=15, out_cap=8)
    arena, jobs = b.arrays()
    assert (i, j) == (0, 1) and jobs[0]["out_cap"] == 7 and jobs[1]["out_off"] % 4 == 0
    assert bytes(arena[jobs[0]["a_off"]:jobs[0]["a_off"] + 4]) == b"ACGT"
