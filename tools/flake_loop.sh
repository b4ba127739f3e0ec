#!/bin/bash
# N runs of est-fact on test_gtf5..8 with the default thread count, in-process and through est-factd;
# every run's five md5s against the committed expectation.  Prints one line per case and form; exit 1 on any mismatch.
N=${1:-25}
cd /root/repo
python - "$N" <<'PY'
import os, sys, subprocess, tempfile
sys.path.insert(0, "tests")
import estfact_util as U
n = int(sys.argv[1]); bad = 0
tmp = tempfile.mkdtemp(prefix="flake_")
srv = U.Server(os.path.join(U.ROOT, "pintron_b200", "bin", "est-factd"), tmp)
try:
    for case in ("test_gtf5", "test_gtf6", "test_gtf7", "test_gtf8"):
        d = os.path.join(tmp, case); os.makedirs(d)
        exp = U.unpack(case, d)
        for form, extra, env in (("inproc", ["--engine", "inproc"], None), ("est-factd", ["--engine", "daemon"], srv.env)):
            miss = 0
            for i in range(n):
                p = subprocess.run([U.GPU_BIN, "--quiet", *extra], cwd=d, env=env, capture_output=True)
                got = U.md5s(d) if p.returncode == 0 else None
                if got != {f: exp[f] for f in U.FILES}:
                    miss += 1
                    print(f"MISMATCH {case} {form} run {i} rc={p.returncode}", flush=True)
            bad += miss
            print(f"{case} {form}: {n - miss}/{n} identical", flush=True)
finally:
    srv.stop()
sys.exit(1 if bad else 0)
PY
