"""Shared helpers for the est-fact level parity tests: unpack a golden case, run a binary on it, compare md5s."""
import hashlib
import json
import lzma
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden", "estfact")
FILES = ["raw-multifasta-out.txt", "processed-ests.txt", "megs.txt", "processed-megs.txt", "meg-edges.txt"]
CPU_BIN = os.path.join(HERE, "_build", "est-fact-oracle-backend")
GPU_BIN = os.path.join(ROOT, "pintron_b200", "bin", "est-fact")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "est-fact")


def cases():
    return sorted(d for d in os.listdir(GOLD) if os.path.exists(os.path.join(GOLD, d, "expected.json")))


def unpack(case, dst):
    for f in ("genomic.txt", "ests.txt"):
        with lzma.open(os.path.join(GOLD, case, f + ".xz")) as i, open(os.path.join(dst, f), "wb") as o:
            o.write(i.read())
    return json.load(open(os.path.join(GOLD, case, "expected.json")))


def run(binary, cwd, *args, timeout=1200, env=None):
    """Runs an est-fact binary.  Unless the caller names one, the engine is in-process: tests must not leave (or pick up)
    a resident est-factd behind their back; the server form has its own tests, on private sockets."""
    if binary != REF_BIN and "--engine" not in args and not (env and "EST_FACT_ENGINE" in env):
        args = (*args, "--engine", "inproc")
    p = subprocess.run([binary, *args], cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=timeout, env=env)
    assert p.returncode == 0, p.stderr.decode("latin1")[-2000:]
    with open(os.path.join(cwd, "stderr.txt"), "w") as f:
        f.write(p.stderr.decode("latin1"))
    return p.stderr.decode("latin1")


def md5s(cwd):
    out = {}
    for f in FILES:
        data = open(os.path.join(cwd, f), "rb").read()
        out[f] = {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data)}
    return out


def check_case(binary, case, tmp_path, *args, env=None):
    exp = unpack(case, str(tmp_path))
    run(binary, str(tmp_path), *args, env=env)
    got = md5s(str(tmp_path))
    for f in FILES:
        if got[f] != exp[f]:
            hint = ""
            full = os.path.join(GOLD, case, f + ".xz")
            if os.path.exists(full):
                want = lzma.open(full).read().split(b"\n")
                have = open(os.path.join(str(tmp_path), f), "rb").read().split(b"\n")
                for n, (a, b) in enumerate(zip(want, have)):
                    if a != b:
                        hint = f" first differing line {n + 1}: expected {a[:120]!r} got {b[:120]!r}"
                        break
            raise AssertionError(f"{case}/{f}: {got[f]} != {exp[f]}{hint}")


def build_cpu_binary():
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    return CPU_BIN


OPTION_SETS = [
    ["-l", "18"], ["-l", "12"], ["--no-transitive-reduction", "--no-short-edge-compaction"], ["--retain-externals=false"],
    ["-B", "60", "--max-intron-length", "5000"], ["-d", "0.5", "-p", "0.4", "-s", "0.4"],
    ["--suff-pref-length-intron", "50", "--suff-pref-length-est", "20", "--suff-pref-length-genomic", "20"],
    ["--max-difference-of-coverage", "0.2", "--max-difference-of-gap-length", "-1", "--complexity-threshold", "10"],
    ["-D", "10", "--max-no-of-factorizations", "2"], ["--max-pairings-in-CMEG", "10", "--max-shortest-pairing-frequence", "0.1"],
]


def check_options_vs_reference(binary, case, tmp_path, opts, *extra):
    """Same flags to the unmodified reference binary and to `binary`; the five output files must be identical."""
    import shutil
    a, b = os.path.join(str(tmp_path), "ref"), os.path.join(str(tmp_path), "ours")
    os.makedirs(a); os.makedirs(b)
    unpack(case, a)
    for f in ("genomic.txt", "ests.txt"):
        shutil.copy(os.path.join(a, f), os.path.join(b, f))
    run(REF_BIN, a, *opts)
    run(binary, b, *opts, *extra)
    assert md5s(a) == md5s(b), (case, opts)


class Server:
    """A private est-factd (own socket in `tmp`) for the tests of the server form; `binary` = the est-factd next to the
    est-fact under test (tests/_build/est-factd for the CPU stand-in, pintron_b200/bin/est-factd on a GPU box)."""

    def __init__(self, daemon_bin, tmp, *args):
        import time
        self.sock = os.path.join(str(tmp), "efd.sock")
        self.log = open(os.path.join(str(tmp), "est-factd.log"), "wb")
        self.proc = subprocess.Popen([daemon_bin, "--socket", self.sock, "--foreground", *args], stdout=self.log, stderr=self.log)
        t0 = time.time()
        while not os.path.exists(self.sock):
            assert self.proc.poll() is None, "est-factd exited: " + open(self.log.name, "rb").read().decode("latin1")[-1500:]
            assert time.time() - t0 < 120, "est-factd did not come up"
            time.sleep(0.02)
        self.env = dict(os.environ, EST_FACTD_SOCKET=self.sock, EST_FACT_NO_SPAWN="1")

    def stop(self):
        if self.proc.poll() is None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=20)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        self.log.close()

    def text(self):
        return open(self.log.name, "rb").read().decode("latin1")
