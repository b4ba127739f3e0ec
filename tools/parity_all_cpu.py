#!/usr/bin/env python
"""Host program + CPU oracle backend (tests/_build/est-fact-oracle-backend) on ALL regression fixtures, compared with the
md5s the unmodified reference produced (tests/golden/estfact/*/expected.json).  Slow cases take minutes on CPU:
    python tools/parity_all_cpu.py [-j 6]  >  profiles/<round>_parity_cpu.txt"""
import concurrent.futures as cf
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import estfact_util as U


def one(case):
    t0 = time.time()
    with tempfile.TemporaryDirectory() as tmp:
        try:
            U.check_case(U.CPU_BIN, case, tmp, "--quiet", "--threads", "2")
            return case, "identical (5 files)", time.time() - t0
        except AssertionError as e:
            return case, "MISMATCH " + str(e)[:200], time.time() - t0
        except Exception as e:      # the CPU oracle backend needs more than check_case's 20 minutes on the largest cases
            return case, "not finished on CPU (" + type(e).__name__ + "); covered by tests/test_estfact_gpu.py on the GPU", time.time() - t0


if __name__ == "__main__":
    U.build_cpu_binary()
    jobs = int(sys.argv[sys.argv.index("-j") + 1]) if "-j" in sys.argv else 4
    cases = sorted(os.listdir(os.path.join(ROOT, "tests", "golden", "estfact")))
    with cf.ThreadPoolExecutor(jobs) as ex:
        for case, verdict, sec in ex.map(one, cases):
            print(f"{case:45s} {verdict}  ({sec:.0f} s)", flush=True)
