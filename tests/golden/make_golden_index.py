#!/usr/bin/env python
"""Regenerates the index-only goldens by running the COMPILED REFERENCE (oracle/_ref, built by `make -C oracle ref`):

  big100.json   sha256 of the six files `vargeno index` writes for the 105 Mbp repeat-family set
  ucscA.json    UCSC snp-table input: sha256 of what the hidden sub-commands `vargeno ucscd` / `ucscbf`
                (src/qv.cc:1954-2008, 2225-2238) write, of `vargeno filt` (src/qv.cc:2009-2025, src/dict_filt.c) run on that
                dictionary with the snp_locs table of `vargeno-b200 index --snp-locs` (the reference writes that table only when
                compiled with GEN_FLT_DATA, src/qv.h:10), and
  ucscA.out.vcf the VCF the reference `geno` writes with the UCSC-built index.

usage: python tests/golden/make_golden_index.py [big100] [ucscA]      (build container only: needs oracle/_ref, ~20 GB RAM)
"""
import hashlib
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from vargeno_b200 import build as vb  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
INDEX_EXT = ["ref.dict", "snp.dict", "ref.bf", "snp.bf", "ref.bf.lite.bf", "chrlens"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def run(cmd, **kw):
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, **kw)


def big100(tmp):
    ds = datasets.make_big100(os.path.join(tmp, "big100"))
    prefix = os.path.join(tmp, "big100", "refix")
    run([orc.REF_BIN, "index", ds.fasta, ds.vcf, prefix])
    man = {"dataset": "big100 (tests/datasets.py make_big100)",
           "made_by": "oracle/_ref/vargeno index (the unmodified reference compiled by oracle/Makefile)",
           "index": {e: sha256(prefix + "." + e) for e in INDEX_EXT},
           "index_bytes": {e: os.path.getsize(prefix + "." + e) for e in INDEX_EXT},
           "inputs": {"ref.fa": sha256(ds.fasta), "snp.vcf": sha256(ds.vcf)}}
    json.dump(man, open(os.path.join(GOLD, "big100.json"), "w"), indent=1)


def ucsc_a(tmp):
    d = os.path.join(tmp, "ucscA")
    ds = datasets.make_ucsc_a(d)
    p = os.path.join(d, "u")
    # the sub-commands take explicit file names and write the chrlens next to the FASTA (src/qv.cc:1963-1976)
    run([orc.REF_BIN, "ucscd", ds.fasta, ds.txt, p + ".ref.dict", p + ".snp.dict"])
    run([orc.REF_BIN, "ucscbf", ds.fasta, ds.txt, p + ".ref.bf", p + ".snp.bf"])
    shutil.copy(ds.fasta + ".chrlens", p + ".chrlens")
    out_vcf = os.path.join(GOLD, "ucscA.out.vcf")
    run([orc.REF_BIN, "geno", p, ds.fastq, ds.vcf, out_vcf])
    # the snp_locs table comes from our host parser (the stock reference cannot write it); `filt` itself is the reference's
    vb.build()
    locs = os.path.join(d, "snp_locs")
    run([vb.HOST_BIN, "index", ds.fasta, ds.txt, os.path.join(d, "unused"), "--dump-parse", os.path.join(d, "dump.txt"), "--snp-locs", locs])
    run([orc.REF_BIN, "filt", p + ".ref.dict", locs, p + ".filt.dict"])
    man = {"dataset": "ucscA (tests/datasets.py make_ucsc_a)",
           "made_by": "oracle/_ref/vargeno ucscd + ucscbf + geno + filt (the unmodified reference compiled by oracle/Makefile)",
           "index": {e: sha256(p + "." + e) for e in INDEX_EXT}, "index_bytes": {e: os.path.getsize(p + "." + e) for e in INDEX_EXT},
           "filt": {"snp_locs": sha256(locs), "filt.dict": sha256(p + ".filt.dict"), "filt_bytes": os.path.getsize(p + ".filt.dict")},
           "inputs": {"ref.fa": sha256(ds.fasta), "snp.txt": sha256(ds.txt), "snp.vcf": sha256(ds.vcf), "reads.fq": sha256(ds.fastq)}}
    json.dump(man, open(os.path.join(GOLD, "ucscA.json"), "w"), indent=1)


if __name__ == "__main__":
    assert orc.have_ref(), "build oracle/_ref first: make -C oracle ref"
    which = sys.argv[1:] or ["big100", "ucscA"]
    with tempfile.TemporaryDirectory(prefix="vg_golden_ix_") as tmp:
        if "big100" in which:
            big100(tmp)
        if "ucscA" in which:
            ucsc_a(tmp)
