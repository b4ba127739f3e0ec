#!/usr/bin/env python
"""Whole-program timing probe (run under gpurun): est-fact on a synthetic workload, in-process engine vs est-factd,
a few thread counts; prints wall clock and the program's own INFO lines.  Not a benchmark of record: bench.py is."""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pintron_b200.synth import Synth, ests_fasta_parallel   # noqa: E402
import estfact_util as U   # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "C3"
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
threads = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0]
forms = sys.argv[4].split(",") if len(sys.argv) > 4 else ["inproc", "daemon"]
work = tempfile.mkdtemp(prefix="probe_")
open(os.path.join(work, "genomic.txt"), "wb").write(Synth(wl, reads=1).genome_fasta())
open(os.path.join(work, "ests.txt"), "wb").write(ests_fasta_parallel(wl, reads, 0, reads, procs=max(1, (os.cpu_count() or 2) - 1)))
print(f"probe: {wl} x {reads} reads in {work}, {os.cpu_count()} cores", flush=True)
exe = os.path.join(ROOT, "pintron_b200", "bin", "est-fact")
variants = sys.argv[5].split(";") if len(sys.argv) > 5 else [""]      # extra est-fact flags, one set per variant; "ENV=val" words go to the environment
KEEP = ("scheduler:", "tail (", "thread-seconds", "lane batches", "engine (", "per-EST code by phase", "@Timer Total", "@Timer IO", "device ms per op", "pc profile", "timeline", "round trips")


def run(tag, args, env):
    for f in os.listdir(work):                       # a fresh run directory, as pintron.py's STEP 2 has: only the two inputs
        if f not in ("genomic.txt", "ests.txt"):
            os.remove(os.path.join(work, f))
    env = dict(env or os.environ)
    for a in [a for a in args if "=" in a and not a.startswith("-")]:
        env[a.split("=", 1)[0]] = a.split("=", 1)[1]
    args = [a for a in args if not ("=" in a and not a.startswith("-"))]
    t0 = time.perf_counter()
    p = subprocess.run([exe, *args], cwd=work, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    dt = time.perf_counter() - t0
    print(f"== {tag}: rc={p.returncode} wall {dt:.3f} s -> {reads / dt:.0f} reads/s", flush=True)
    for l in p.stderr.decode("latin1").splitlines():
        if any(k in l for k in KEEP) or p.returncode:
            print("   " + l[:600], flush=True)
    return subprocess.run("md5sum raw-multifasta-out.txt processed-ests.txt | cut -c1-12", shell=True, cwd=work, capture_output=True, text=True).stdout.split()


md5 = {}
for th in threads:
    targs = ["--threads", str(th)] if th else []
    if "inproc" in forms:
        md5[("inproc", th)] = run(f"inproc threads={th or 'default'}", ["--engine", "inproc", *targs], None)
    if "daemon" in forms:
        srv_dir = tempfile.mkdtemp(prefix="probe_srv_")
        t0 = time.perf_counter()
        srv = U.Server(os.path.join(ROOT, "pintron_b200", "bin", "est-factd"), srv_dir)
        print(f"est-factd up in {time.perf_counter() - t0:.3f} s", flush=True)
        try:
            for var in variants:
                for rep in range(int(os.environ.get("PROBE_REPS", "3"))):
                    md5[("daemon", th, var, rep)] = run(f"est-factd threads={th or 'default'} [{var}] run {rep}", ["--engine", "daemon", *targs, *var.split()], srv.env)
            if os.environ.get("PROBE_PROFILE"):
                run(f"est-factd threads={th or 'default'} PC_PROFILE", ["--engine", "daemon", *targs], dict(srv.env, PC_PROFILE="1"))
        finally:
            srv.stop()
        nl = 30 if os.environ.get("PROBE_PROFILE") or os.environ.get("PC_PROFILE_HOST") else 4
        print("   server log tail: " + "\n      ".join(l[:700] for l in srv.text().splitlines()[-nl:]), flush=True)
print("md5s identical across runs:", len(set(map(tuple, md5.values()))) == 1, flush=True)
