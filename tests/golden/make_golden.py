#!/usr/bin/env python
"""Regenerates tests/golden/* by running the COMPILED REFERENCE (oracle/_ref, built by `make -C oracle ref`
from /root/reference) on the synthetic data sets of tests/datasets.py.

Only runs in the build container (needs oracle/_ref and ~20 GB RAM for the reference's 16 GiB jumpgate).
What is committed per data set <name>:
  <name>.json     sha256 of the generated inputs and of the five index files the reference `index` wrote
                  (+ the unused .lite.bf), record counts, reference run time
  <name>.npz      per-read vote trace of the reference `geno` (ord, flags, target, freq, n_ref, n_snp,
                  digest of every recorded hit context), the pileup dump, and the calls with %.17g confidence
  <name>.out.vcf  the VCF the reference wrote

usage: python tests/golden/make_golden.py [name ...]
"""
import hashlib
import json
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import datasets  # noqa: E402
from oracle import oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
INDEX_EXT = ["ref.dict", "snp.dict", "ref.bf", "snp.bf", "ref.bf.lite.bf", "chrlens"]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def mix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def ctx_digest(list_id, position, kmer_pos, kmer, mod):
    """numpy twin of vgo_ctx_digest (oracle/vg_oracle.c)."""
    with np.errstate(over="ignore"):
        h = mix64(np.asarray(kmer, np.uint64) + np.uint64(0x9E3779B97F4A7C15))
        h = mix64(h ^ ((np.asarray(position, np.uint64) << np.uint64(32)) | np.asarray(kmer_pos, np.uint64)))
        h = mix64(h ^ ((np.asarray(mod, np.uint64) << np.uint64(8)) | np.asarray(list_id, np.uint64)))
    return h


def parse_trace(path, n_reads):
    res = np.zeros(n_reads, dtype=orc.READ_RESULT)
    res["flags"] = orc.F_SKIPPED          # reads that never reach the trace point were skipped (N)
    cur = -1
    acc = {}
    with open(path) as f:
        for line in f:
            t = line.split()
            if t[0] == "R":
                cur = int(t[1])
                fl = (orc.F_REVCOMPL if t[2] == "1" else 0) | (orc.F_PROCESS if t[3] == "1" else 0) | \
                     (orc.F_AMBIGUOUS if t[6] == "1" else 0)
                freq = int(t[5])
                # best != NULL <=> freq printed non-zero (freq of a best entry is >= 2)
                if freq:
                    fl |= orc.F_HASBEST
                res[cur] = (fl, int(t[4]), freq, int(t[7]), int(t[8]), 0, 0)
                acc[cur] = []
            else:
                acc[cur].append((0 if t[0] == "r" else 1, int(t[1]), int(t[2]), int(t[3]), int(t[4])))
    for k, lst in acc.items():
        if lst:
            a = np.array(lst, dtype=np.uint64)
            with np.errstate(over="ignore"):
                res["ctx_hash"][k] = np.sum(ctx_digest(a[:, 0], a[:, 1], a[:, 2], a[:, 3], a[:, 4]), dtype=np.uint64)
    return res


def make(name):
    assert orc.have_ref(), "build the reference first: make -C oracle ref"
    d = tempfile.mkdtemp(prefix="vg_gold_" + name + "_")
    try:
        ds = datasets.MAKERS[name](d)
        prefix = os.path.join(d, "ref")
        orc.run_ref_index(ds.fasta, ds.vcf, prefix)
        out_vcf = os.path.join(d, "out.vcf")
        secs = orc.run_ref_geno(prefix, ds.fastq, ds.vcf, out_vcf, trace=os.path.join(d, "trace"), dump=os.path.join(d, "dump"))
        res = parse_trace(os.path.join(d, "trace"), ds.n_reads)
        # the trace has no 'passes' column: leave 0
        sites, calls = orc.parse_dump(os.path.join(d, "dump"))
        calls = [c for c in calls if c[2] != 0]
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), reads=res, sites=sites,
                            call_chr=np.array([c[0] for c in calls]), call_pos=np.array([c[1] for c in calls], np.int64),
                            call_gt=np.array([c[2] for c in calls], np.uint8), call_conf=np.array([c[3] for c in calls], np.float64))
        shutil.copy(out_vcf, os.path.join(GOLD, name + ".out.vcf"))
        man = {
            "dataset": name, "n_reads": ds.n_reads,
            "inputs": {k: sha256(p) for k, p in (("fasta", ds.fasta), ("vcf", ds.vcf), ("fastq", ds.fastq))},
            "index": {e: sha256(prefix + "." + e) for e in INDEX_EXT},
            "index_bytes": {e: os.path.getsize(prefix + "." + e) for e in INDEX_EXT},
            "reference_geno_time_s": secs,
            "placed_reads": int(np.count_nonzero(res["flags"] & orc.F_PROCESS)),
            "skipped_reads": int(np.count_nonzero(res["flags"] & orc.F_SKIPPED)),
            "sites": int(sites.size), "calls": len(calls),
            "made_by": "tests/golden/make_golden.py with oracle/_ref/vargeno{,_instr} (reference qv.cc sha256 in oracle/Makefile)",
        }
        with open(os.path.join(GOLD, name + ".json"), "w") as f:
            json.dump(man, f, indent=1, sort_keys=True)
        print(name, json.dumps({k: man[k] for k in ("n_reads", "placed_reads", "skipped_reads", "sites", "calls")}))
    finally:
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    for n in (sys.argv[1:] or list(datasets.MAKERS)):
        make(n)
