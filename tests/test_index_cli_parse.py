"""Host half of `vargeno-b200 index` (csrc/host/index_host.cpp): its FASTA readers and its two VCF walks must hand the device
builder exactly what tools/index_builder (pinned to the reference-built files by tests/test_index_builder.py) derives from the
same text -- contig geometry, the dictionary-side SNP lines (src/dictgen.c:599-748) and the Bloom-filter-side lines
(src/generate_bf.cc:203-246).  No GPU involved: `--dump-parse` stops before the device step."""
import os
import subprocess

import numpy as np
import pytest

from vargeno_b200 import build as vb
from vargeno_b200.tools import index_builder as ib


def _expected(ds):
    raw_names, raw_seqs = ib.read_fasta_raw(ds.fasta)
    names, seqs = ib.normalise_fasta(raw_names, raw_seqs)
    starts = np.concatenate([[0], np.cumsum([s.size for s in seqs])[:-1]]).astype(np.int64)
    lines = ib.parse_vcf_for_dict(ds.vcf, names, seqs)
    bc, bp = ib.bf_line_positions(ds.vcf, raw_names, raw_seqs)
    out = ["C %s %d %d" % (n, s, q.size) for n, s, q in zip(names, starts, seqs)]
    out += ["D %d %d %d %d" % (starts[c] + p, int(r) | (int(a) << 2), rf, af)
            for c, p, r, a, rf, af in zip(lines.contig, lines.index0, lines.ref_code, lines.alt_code, lines.ref_freq, lines.alt_freq)]
    out += ["B %d" % (starts[c] + p) for c, p in zip(bc, bp)]
    return out


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_cli_parse_matches_numpy_builder(cache, name, tmp_path):
    vb.build()
    ds = cache.dataset(name)
    dump = str(tmp_path / "parse.txt")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.vcf, str(tmp_path / "unused"), "--dump-parse", dump],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    got = open(dump).read().splitlines()
    want = _expected(ds)
    assert len(got) == len(want)
    assert got == want


def test_cli_parse_quirks(tmp_path):
    """Hand-written inputs for the corners the synthetic sets do not reach: lower-case FASTA, header with description and
    '|', a VCF without the chr prefix, CAF position carried over between records, multi-base alleles, REF == ALT, SNPs too
    close to a contig end or next to N, an unknown contig (stale-sequence rule on the Bloom-filter side)."""
    vb.build()
    rng = np.random.default_rng(5)
    a = "".join("ACGT"[i] for i in rng.integers(0, 4, 400))
    b = "".join("acgt"[i] for i in rng.integers(0, 4, 300))
    b = b[:100] + "NNNN" + b[104:]
    fasta = tmp_path / "q.fa"
    fasta.write_text(">chr1 first contig|x\n" + "\n".join(a[i:i + 60] for i in range(0, 400, 60)) + "\n>chr2\n" +
                     "\n".join(b[i:i + 70] for i in range(0, 300, 70)) + "\n")

    def alt(c):
        return "ACGT"[("ACGT".index(c.upper()) + 1) % 4]
    recs = []
    for chrom, seq, poss in (("1", a, [10, 40, 41, 100, 200, 369, 380]), ("chr2", b, [50, 90, 110, 130, 150, 250]), ("chr9", b, [60])):
        for p1 in poss:
            r = seq[p1 - 1] if p1 % 20 else seq[p1 - 1].upper()      # chr2 is lower case: the Bloom-filter side compares REF with the raw character
            recs.append("%s\t%d\trs%d\t%s\t%s\t.\t.\tRS=1;CAF=0.%d,0.%d;X=2" % (chrom, p1, p1, r, alt(r), p1 % 10, (p1 * 7) % 10))
    recs.insert(3, "1\t120\trsm\t%s\t%s,%s\t.\t.\tCAF=0.5,0.5" % (a[119], alt(a[119]), a[119]))      # multi-ALT
    recs.insert(5, "1\t130\trsi\t%sA\t%s\t.\t.\tCAF=0.5,0.5" % (a[129], a[129]))                     # deletion
    recs.insert(6, "1\t140\trse\t%s\t%s\t.\t.\tCAF=0.5,0.5" % (a[139], a[139]))                      # REF == ALT
    recs.insert(7, "1\t150\trsn\t%s\t%s\t.\t.\tA=1;B=2;C=3;0.25,0.75" % (a[149], alt(a[149])))       # no CAF key: token index carried over
    vcf = tmp_path / "q.vcf"
    vcf.write_text("##fileformat=VCFv4.0\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n" + "\n".join(recs) + "\n")

    class DS:
        pass
    ds = DS()
    ds.fasta, ds.vcf = str(fasta), str(vcf)
    want = _expected(ds)
    assert sum(l.startswith("D ") for l in want) >= 8 and sum(l.startswith("B ") for l in want) >= 2
    dump = str(tmp_path / "parse.txt")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.vcf, str(tmp_path / "unused"), "--dump-parse", dump],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(dump).read().splitlines() == want


def test_cli_index_usage(tmp_path):
    vb.build()
    p = subprocess.run([vb.HOST_BIN, "index", "only-one-arg"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0
    p = subprocess.run([vb.HOST_BIN, "index", str(tmp_path / "missing.fa"), str(tmp_path / "missing.vcf"), str(tmp_path / "p")],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0
