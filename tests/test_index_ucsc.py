"""SURVEY 8(f)-4 without a GPU: the UCSC snp-table input of `vargeno-b200 index` (host parser, csrc/host/index_host.cpp) and
`vargeno-b200 filt`, against what the COMPILED REFERENCE's hidden sub-commands wrote for the same text (tests/golden/ucscA.json,
made by tests/golden/make_golden_index.py from `vargeno ucscd` / `ucscbf` / `filt`: src/qv.cc:1954-2025, 2225-2238;
src/dictgen.c:350-540; src/generate_bf.cc:439-592; src/dict_filt.c:23-79).

The host parser's output (--dump-parse) goes through the numpy dictionary builder (pinned to reference-built files by
tests/test_index_builder.py) and a literal restatement of the 33-values-per-record SNP filter; the bytes must hash to the
reference's."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from vargeno_b200 import build as vb
from vargeno_b200.tools import index_builder as ib

GOLD = os.path.join(os.path.dirname(__file__), "golden")
U64 = np.uint64


def _sha_bytes(*parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(memoryview(p))
    return h.hexdigest()


@pytest.fixture(scope="module")
def parsed(cache, tmp_path_factory):
    vb.build()
    ds = cache.dataset("ucscA")
    d = tmp_path_factory.mktemp("ucsc")
    dump, locs = str(d / "dump.txt"), str(d / "snp_locs")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.txt, str(d / "unused"), "--dump-parse", dump, "--snp-locs", locs],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    raw_names, raw_seqs = ib.read_fasta_raw(ds.fasta)
    names, seqs = ib.normalise_fasta(raw_names, raw_seqs)
    rows = [l.split() for l in open(dump)]
    return {"ds": ds, "dir": d, "locs": locs, "names": names, "seqs": seqs, "rows": rows,
            "man": json.load(open(os.path.join(GOLD, "ucscA.json")))}


def test_inputs_are_the_ones_the_goldens_were_made_from(parsed):
    ds, man = parsed["ds"], parsed["man"]
    for name, path in (("ref.fa", ds.fasta), ("snp.txt", ds.txt), ("reads.fq", ds.fastq)):
        assert _sha_bytes(open(path, "rb").read()) == man["inputs"][name], name


def test_ucsc_dictionaries_match_the_reference(parsed):
    """make_snp_dict (src/dictgen.c:350-540): strand handling, allele / frequency swap, first usable allele of `observed`,
    all the record kinds the filters drop."""
    seqs, man = parsed["seqs"], parsed["man"]
    starts = np.concatenate([[0], np.cumsum([s.size for s in seqs])[:-1]]).astype(np.int64)
    D = [r for r in parsed["rows"] if r[0] == "D"]
    assert len(D) > 1000
    g = np.array([int(r[1]) for r in D], np.int64)
    code = np.array([int(r[2]) for r in D], np.uint8)
    ci = (np.searchsorted(starts, g, side="right") - 1).astype(np.int32)
    lines = ib.SnpLines(ci, g - starts[ci], code & 3, code >> 2, np.array([int(r[3]) for r in D], np.uint8), np.array([int(r[4]) for r in D], np.uint8))
    ref, ref_aux, pck = ib.build_ref_dict(seqs, want_kmers=True)
    snp, snp_aux = ib.build_snp_dict(lines, seqs, pck)
    assert _sha_bytes(np.array([snp.size, snp_aux.size], dtype="<u8").tobytes(), snp.tobytes(), snp_aux.tobytes()) == man["index"]["snp.dict"]
    assert _sha_bytes(np.array([ref.size, ref_aux.shape[0]], dtype="<u8").tobytes(), ref.tobytes(),
                      np.ascontiguousarray(ref_aux, dtype="<u4").tobytes()) == man["index"]["ref.dict"]
    # both strands and both allele orders are present, and records were dropped
    n_rec = sum(1 for l in open(parsed["ds"].txt) if not l.startswith("#"))
    assert len(D) < n_rec


def test_ucsc_snp_filter_matches_the_reference(parsed):
    """constructBfFromUcsc (src/generate_bf.cc:538-556): LO40 of the 32-mer in front of the SNP (0 when it holds an N) and of the
    32 alternative-allele k-mers up to the first N."""
    seqs, man = parsed["seqs"], parsed["man"]
    cat = np.concatenate(seqs)
    vals = []
    U = [r for r in parsed["rows"] if r[0] == "U"]
    assert len(U) > 1000
    for r in U:
        p, alt = int(r[1]), int(r[2])
        w = ib._CODE[cat[p - 32:p]]
        if np.any(w > 3):
            vals.append(0)
            continue
        km = 0
        for j in range(32):
            km |= int(w[j]) << (2 * j)
        vals.append(km & 0xFFFFFFFFFF)
        for j in range(32):
            c = alt if j == 0 else int(ib._CODE[cat[p + j]])
            if c > 3:
                break
            km = (km >> 2) | (c << 62)
            vals.append(km & 0xFFFFFFFFFF)
    words = np.zeros((ib.SNP_BF_BITS + 63) // 64, dtype=U64)
    ib._set_bits(words, ib.hash40(np.array(vals, dtype=U64)) % U64(ib.SNP_BF_BITS))
    assert _sha_bytes(np.array([ib.SNP_BF_BITS], dtype="<u8").tobytes(), words.tobytes()) == man["index"]["snp.bf"]


def test_filt_matches_the_reference(parsed, tmp_path):
    """dict_filt (src/dict_filt.c:23-79) on the reference dictionary of this set, with the snp_locs table `--snp-locs` wrote."""
    seqs, man = parsed["seqs"], parsed["man"]
    assert _sha_bytes(open(parsed["locs"], "rb").read()) == man["filt"]["snp_locs"]
    ref, ref_aux, _ = ib.build_ref_dict(seqs, want_kmers=True)
    rd = str(tmp_path / "ref.dict")
    with open(rd, "wb") as f:
        f.write(np.array([ref.size, ref_aux.shape[0]], dtype="<u8").tobytes())
        ref.tofile(f)
        np.ascontiguousarray(ref_aux, dtype="<u4").tofile(f)
    out = str(tmp_path / "filt.dict")
    p = subprocess.run([vb.HOST_BIN, "filt", rd, parsed["locs"], out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert os.path.getsize(out) == man["filt"]["filt_bytes"]
    assert _sha_bytes(open(out, "rb").read()) == man["filt"]["filt.dict"]
    assert "New size:" in p.stdout and "Removed:" in p.stdout


def test_ucsc_mismatch_is_loud(parsed, tmp_path):
    """A record whose reference base disagrees with the FASTA makes the reference exit (src/dictgen.c:431-437): an error here."""
    ds = parsed["ds"]
    lines = open(ds.txt).read().splitlines()
    k = next(i for i, l in enumerate(lines) if not l.startswith("#"))
    f = lines[k].split("\t")
    f[7] = f[8] = "ACGT"[("ACGT".index(f[7]) + 1) % 4]
    lines[k] = "\t".join(f)
    bad = tmp_path / "bad.txt"
    bad.write_text("\n".join(lines) + "\n")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, str(bad), str(tmp_path / "x"), "--dump-parse", str(tmp_path / "d")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode != 0 and "Mismatch" in p.stderr
