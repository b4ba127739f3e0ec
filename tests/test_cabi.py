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


def declared_symbols(header="pintron_cuda.h"):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = text.split("client-side helpers")[0]          # the header-only lane protocol functions are not exports
    return sorted(set(re.findall(r"\b(pc_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(lib):
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pintron_cuda.h but not exported"
    assert set(binding.EXPORTS) <= set(syms)


def test_exports_every_engine_symbol(lib):
    """include/pintron_engine.h: the batch engine + pc_submit_parts."""
    syms = declared_symbols("pintron_engine.h")
    assert len(syms) >= 12, syms
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/pintron_engine.h but not exported"
    assert set(binding.ENGINE_EXPORTS) == set(syms)
    assert lib.pc_engine_backend() == b"cuda-sm100a"


def test_lane_struct_layout():
    """The shared-memory lane table is read by C (host, engine) and by Python (tests, bench): same layout."""
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "pintron_engine.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n", sizeof(pce_lane), ' \
          'offsetof(pce_hdr, lanes), offsetof(pce_hdr, doorbell), offsetof(pce_lane, arena_off), offsetof(pce_lane, var_cap));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True).stdout.split()
    assert [int(x) for x in out] == [128, binding.PCE_LANES_OFFSET, 8, binding.pce_lane.arena_off.offset, binding.pce_lane.var_cap.offset]


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
