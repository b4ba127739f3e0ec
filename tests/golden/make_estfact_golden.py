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
FULL = {"test-AMBN", "test-788", "test-mattia1", "test-mattia3", "edge-cases"}


def edge_case_inputs(tmp):
    """Edge inputs the reference handles (an EMPTY sequence record makes it segfault, so that one is not a fixture):
    ESTs shorter than the word length, all-N, pure polyA, lower case, IUPAC letters + RefSeq id (fixed strand), a
    header without any field, /fixed_strand=1, and a verbatim piece of the genome; on the genome of test-mattia3."""
    src = os.path.join(REF, "regressionTest", "test-mattia3")
    g = open(os.path.join(src, "genomic.txt"), "rb").read()
    recs = open(os.path.join(src, "ests.txt"), "rb").read().split(b">")[1:]
    body = lambda r: r.split(b"\n", 1)[1]
    ests = (b">tiny /gb=T1 /clone_end=3'\nACGTA\n" + b">l14 /gb=T1b /clone_end=3'\nACGTACGTACGTAC\n" +
            b">l20 /gb=T1c /clone_end=3'\nACGTACGTACGTACGGATCA\n" + b">allN /gb=T2 /clone_end=5'\n" + b"N" * 80 + b"\n" +
            b">polyA /gb=T3 /clone_end=3'\n" + b"A" * 120 + b"\n" + b">" + recs[0] +
            b">lower /gb=T5 /clone_end=3'\n" + body(recs[1]).lower() + b">iupac /gb=NM_T6\n" + body(recs[2]).replace(b"A", b"R", 3) +
            b">plainheader\n" + body(recs[4]) + b">x /gb=F1 /clone_end=5' /fixed_strand=1\n" + body(recs[5]) +
            b">gp /gb=G1 /clone_end=3'\n" + g.split(b"\n", 1)[1].replace(b"\n", b"").replace(b"\r", b"")[1000:1400] + b"\n" + b">" + recs[3])
    open(os.path.join(tmp, "genomic.txt"), "wb").write(g)
    open(os.path.join(tmp, "ests.txt"), "wb").write(ests)


def main():
    out_root = os.path.join(HERE, "estfact")
    cases = sys.argv[1:] or (sorted(d for d in os.listdir(os.path.join(REF, "regressionTest"))
                                    if os.path.exists(os.path.join(REF, "regressionTest", d, "ests.txt"))) + ["edge-cases"])
    for case in cases:
        src = os.path.join(REF, "regressionTest", case)
        tmp = tempfile.mkdtemp(prefix="golden_")
        if case == "edge-cases":
            edge_case_inputs(tmp)
        for f in ("genomic.txt", "ests.txt"):
            if case != "edge-cases":
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
    # the reference's own pipeline-level golden for test-AMBN (regressionTest/test-AMBN/referenceOutput/full.json, older key
    # names): the 13 predicted introns as (relative start, relative end, supporting ESTs) — tests/test_pipeline.py
    full = json.load(open(os.path.join(REF, "regressionTest", "test-AMBN", "referenceOutput", "full.json")))
    introns = sorted((v["relative start"], v["relative end"], v["number supporting EST"]) for v in full["introns"].values())
    json.dump(introns, open(os.path.join(out_root, "test-AMBN", "golden_introns.json"), "w"))


if __name__ == "__main__":
    main()
