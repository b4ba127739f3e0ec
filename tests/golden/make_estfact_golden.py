#!/usr/bin/env python
"""Generates tests/golden/estfact/: est-fact level fixtures produced by the UNMODIFIED reference (oracle/_ref/est-fact,
built by `make -C oracle ref`) on the reference's own regression inputs ($REF/regressionTest/<case>/{genomic,ests}.txt).

  <case>/genomic.txt.xz, ests.txt.xz     the inputs (test data of the reference, compressed)
  <case>/expected.json                   md5 + size of raw-multifasta-out.txt, processed-ests.txt, megs.txt,
                                         processed-megs.txt, meg-edges.txt as the reference wrote them
  <case>/raw-multifasta-out.txt.xz       (small cases only) the full expected bytes, for readable diffs

Run in the container that has /root/reference:  python tests/golden/make_estfact_golden.py
"""
import hashlib
import json
import lzma
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("PINTRON_REF", "/root/reference")
EXE = os.path.join(ROOT, "oracle", "_ref", "est-fact")
FILES = ["raw-multifasta-out.txt", "processed-ests.txt", "megs.txt", "processed-megs.txt", "meg-edges.txt"]
FULL = {"test-AMBN", "test-788", "test-mattia1", "test-mattia3"}


def main():
    out_root = os.path.join(HERE, "estfact")
    cases = sys.argv[1:] or sorted(d for d in os.listdir(os.path.join(REF, "regressionTest"))
                                   if os.path.exists(os.path.join(REF, "regressionTest", d, "ests.txt")))
    for case in cases:
        src = os.path.join(REF, "regressionTest", case)
        tmp = tempfile.mkdtemp(prefix="golden_")
        for f in ("genomic.txt", "ests.txt"):
            shutil.copy(os.path.join(src, f), tmp)
        subprocess.run([EXE], cwd=tmp, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        dst = os.path.join(out_root, case)
        os.makedirs(dst, exist_ok=True)
        for f in ("genomic.txt", "ests.txt"):
            with lzma.open(os.path.join(dst, f + ".xz"), "wb", preset=9) as o:
                o.write(open(os.path.join(tmp, f), "rb").read())
        exp = {}
        for f in FILES:
            data = open(os.path.join(tmp, f), "rb").read()
            exp[f] = {"md5": hashlib.md5(data).hexdigest(), "bytes": len(data)}
            if case in FULL and f == "raw-multifasta-out.txt":
                with lzma.open(os.path.join(dst, f + ".xz"), "wb", preset=9) as o:
                    o.write(data)
        exp["n_ests"] = open(os.path.join(tmp, "ests.txt"), "rb").read().count(b">")
        json.dump(exp, open(os.path.join(dst, "expected.json"), "w"), indent=1, sort_keys=True)
        shutil.rmtree(tmp)
        print(case, exp["raw-multifasta-out.txt"])


if __name__ == "__main__":
    main()
