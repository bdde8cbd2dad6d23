"""Named synthetic data sets used by the test-suite and by tests/golden/make_golden.py.

Shapes follow SURVEY.md section 4 / 8(d): S0 smoke, adversarial set A (repeats, N runs, clustered SNPs,
second-allele and indel VCF lines, mixed read lengths, lower case, N in reads, both strands) and
adversarial set B (ref HI32 block >= 100 -> "big" neighbour mode, poly-T = last jumpgate block,
>10-copy repeats -> POS_AMBIGUOUS, SNPs inside repeats -> SNP aux rows, targeted low-quality reads).
Everything is a pure function of the seeds below.
"""
from __future__ import annotations

import dataclasses
import os
from typing import Dict, List

import numpy as np

from vargeno_b200.tools import synth


@dataclasses.dataclass
class Dataset:
    name: str
    dir: str
    fasta: str
    vcf: str
    fastq: str
    n_reads: int
    txt: str = ""      # UCSC snp-table dump of the same SNP list (legacy input format), when the set has one


def _snpset_from_rows(rows) -> synth.SnpSet:
    rows = sorted(set(rows))
    # one record per (contig, pos)
    seen, out = set(), []
    for r in rows:
        if (r[0], r[1]) in seen:
            continue
        seen.add((r[0], r[1]))
        out.append(r)
    a = np.array
    return synth.SnpSet(a([r[0] for r in out], np.int32), a([r[1] for r in out], np.int64),
                        a([r[2] for r in out], np.uint8), a([r[3] for r in out], np.uint8),
                        a([r[4] for r in out], np.float64), a([r[5] for r in out], np.uint8))


def _rows(s: synth.SnpSet):
    return [(int(s.contig[i]), int(s.pos0[i]), int(s.ref[i]), int(s.alt[i]), float(s.caf_ref[i]), int(s.gt[i]))
            for i in range(s.pos0.size)]


def _write(d: str, name: str, g, snps, fq_parts: List[np.ndarray], **vcf_kw) -> Dataset:
    os.makedirs(d, exist_ok=True)
    fa, vcf, fq = (os.path.join(d, x) for x in ("ref.fa", "snp.vcf", "reads.fq"))
    synth.write_fasta(g, fa)
    synth.write_vcf(g, snps, vcf, **vcf_kw)
    text = np.concatenate(fq_parts)
    text.tofile(fq)
    return Dataset(name, d, fa, vcf, fq, int(np.count_nonzero(text == 10)) // 4)


def make_s0(d: str, n_reads: int = 20000) -> Dataset:
    g = synth.make_genome([("chr22", 300000), ("chrX", 150000)], seed=1, n_runs=[(0, 120000, 100), (1, 5000, 37)],
                          repeats=[(400, 3, 1, 1), (200, 12, 0, 2), (64, 5, 2, 3)])
    snps = synth.make_snps(g, 300, seed=1, cluster_frac=0.2)
    haps = synth.donor_haplotypes(g, snps, seed=1)
    fq = synth.simulate_reads(g, haps, n_reads, 150, seed=1)
    return _write(d, "s0", g, snps, [fq])


def _snps_in_repeats(g, seed: int, families, per_family: int, rows):
    """Same SNP (same ALT) at the same offset of every copy of a repeat family: identical alt-allele k-mers from
    several records -> ambiguous SNP-dictionary entries (aux rows; > 10 copies -> POS_AMBIGUOUS)."""
    for fam in families:
        copies = [c for c in g.repeat_copies if c[0] == fam]
        if not copies:
            continue
        flen = copies[0][4]
        for j in range(per_family):
            r = int(synth.rnd64(seed, 40, fam, j))
            off = 40 + r % max(1, flen - 80)
            alt_shift = 1 + (r >> 33) % 3
            for (_, _, ci, o, _) in copies:
                p = o + off
                if p < 32 or p + 32 > g.seqs[ci].size:
                    continue
                ref = int(g.seqs[ci][p])
                if ref == ord("N"):
                    continue
                alt = int(synth.BASES[(int(synth._CODE[ref]) + alt_shift) & 3])
                rows.append((ci, p, ref, alt, 0.3, (r >> 50) % 3))


def make_adv_a(d: str) -> Dataset:
    fams = []
    for f in range(40):
        r = int(synth.rnd64(11, 50, f))
        fams.append((100 + r % 500, 1 + (r >> 20) % 12, (r >> 30) % 3, f + 1))
    g = synth.make_genome([("chr1", 260000), ("chr2", 180000)], seed=11,
                          n_runs=[(0, 50000, 250), (0, 200000, 31), (1, 90000, 1000), (1, 100, 5)], repeats=fams)
    base = synth.make_snps(g, 1000, seed=11, cluster_frac=0.3)
    rows = _rows(base)
    _snps_in_repeats(g, 11, [f[3] for f in fams if f[2] == 0][:8], 2, rows)
    snps = _snpset_from_rows(rows)
    # extra VCF lines: second allele at the same POS, an indel at a SNP POS, a multi-ALT line, an 'N' ALT
    extra: Dict[int, List[str]] = {}
    for i in range(0, snps.pos0.size, 37):
        name = g.names[int(snps.contig[i])]
        pos1 = int(snps.pos0[i]) + 1
        ref = chr(snps.ref[i])
        other = [b for b in "ACGT" if b != ref and b != chr(snps.alt[i])]
        kind = (i // 37) % 4
        if kind == 0:
            extra[i] = ["%s\t%d\trsX%d\t%s\t%s\t.\t.\tRS=1;CAF=0.6,0.4;COMMON=1" % (name, pos1, i, ref, other[0])]
        elif kind == 1:
            nxt = chr(g.seqs[int(snps.contig[i])][int(snps.pos0[i]) + 1])
            extra[i] = ["%s\t%d\trsI%d\t%s%s\t%s\t.\t.\tRS=1;CAF=0.7,0.3;COMMON=1" % (name, pos1, i, ref, nxt, ref)]
        elif kind == 2:
            extra[i] = ["%s\t%d\trsM%d\t%s\t%s,%s\t.\t.\tRS=1;CAF=0.5,0.25,0.25;COMMON=1" % (name, pos1, i, ref, other[0], other[1])]
        else:
            extra[i] = ["%s\t%d\trsN%d\t%s\tN\t.\t.\tRS=1;CAF=0.5,0.5;COMMON=1" % (name, pos1, i, ref)]
    haps = synth.donor_haplotypes(g, snps, seed=11)
    parts, first = [], 0
    for L, n, sub in ((31, 500, 0.0), (64, 4000, 0.01), (100, 6000, 0.02), (150, 14000, 0.005), (250, 5000, 0.015)):
        parts.append(synth.simulate_reads(g, haps, n, L, seed=11 + L, sub_rate=sub, first_id=first, lower_rate=0.1,
                                          n_rate=0.0003, qual_alphabet=b"#5:AFJ"))
        first += n
    return _write(d, "advA", g, snps, parts, extra_lines=extra)


def adv_b_genome():
    """The advB genome with its bookkeeping (repeat copies, motif sites): a pure function of its seeds."""
    motif = b"ACGGTCATGCAAGTCC"
    fams = [(300, 12, 0, 1), (500, 14, 0, 2), (250, 6, 0, 3), (180, 3, 1, 4), (90, 11, 0, 5)]
    return synth.make_genome([("chr7", 400000)], seed=23, n_runs=[(0, 170000, 64)], repeats=fams,
                             motifs=[(motif, 160), (b"T" * 16, 30), (b"T" * 40, 6), (b"A" * 40, 4)])


def make_adv_b(d: str) -> Dataset:
    g = adv_b_genome()
    base = synth.make_snps(g, 400, seed=23, cluster_frac=0.25)
    rows = _rows(base)
    _snps_in_repeats(g, 23, [1, 2, 3, 4, 5], 3, rows)
    # SNPs right next to planted motifs (neighbour veto paths)
    for k, (mi, ci, o) in enumerate(g.motif_sites[:120]):
        p = o + (k % 48) - 16
        if 32 <= p and p + 32 <= g.seqs[ci].size and g.seqs[ci][p] != ord("N"):
            ref = int(g.seqs[ci][p])
            rows.append((ci, p, ref, int(synth.BASES[(int(synth._CODE[ref]) + 1 + k % 3) & 3]), 0.1, k % 3))
    snps = _snpset_from_rows(rows)
    haps = synth.donor_haplotypes(g, snps, seed=23)
    parts = [synth.simulate_reads(g, haps, 8000, 150, seed=23, sub_rate=0.01, lowq_prob=0.5)]
    first = 8000
    # targeted reads: motif on bases 16..31 of k-mer j of the read, all leading qualities low
    starts, revs = [], []
    for k, (mi, ci, o) in enumerate(g.motif_sites):
        for j in range(4):
            s = o - 16 - 32 * j
            if s < 0 or s + 150 > g.seqs[ci].size:
                continue
            starts.append(s)
            revs.append((k + j) % 5 == 0)
    starts = np.array(starts, dtype=np.int64)
    parts.append(synth.simulate_reads(g, haps, starts.size, 150, seed=29, sub_rate=0.02, lowq_prob=1.0, first_id=first,
                                      forced_starts=starts, forced_rev=np.array(revs)))
    first += starts.size
    # reads over the repeat copies
    rs = np.array([o + (k * 7) % max(1, l - 150) for k, (_, _, ci, o, l) in enumerate(g.repeat_copies) if l >= 150] * 20,
                  dtype=np.int64)
    parts.append(synth.simulate_reads(g, haps, rs.size, 150, seed=31, sub_rate=0.01, lowq_prob=0.6, first_id=first,
                                      forced_starts=rs))
    return _write(d, "advB", g, snps, parts, sample_columns=True, declare_gt=True)


def make_ucsc_a(d: str) -> Dataset:
    """The advA genome with its SNP list as a UCSC snp141Common-style table (src/dictgen.c:350-540,
    src/generate_bf.cc:439-592): both strands, shuffled allele order, records of every kind the filters drop."""
    fams = []
    for f in range(40):
        r = int(synth.rnd64(11, 50, f))
        fams.append((100 + r % 500, 1 + (r >> 20) % 12, (r >> 30) % 3, f + 1))
    g = synth.make_genome([("chr1", 260000), ("chr2", 180000)], seed=11,
                          n_runs=[(0, 50000, 250), (0, 200000, 31), (1, 90000, 1000), (1, 100, 5)], repeats=fams)
    rows = _rows(synth.make_snps(g, 1500, seed=12, cluster_frac=0.3))
    _snps_in_repeats(g, 12, [f[3] for f in fams if f[2] == 0][:8], 2, rows)
    snps = _snpset_from_rows(rows)
    os.makedirs(d, exist_ok=True)
    fa, vcf, fq, txt = (os.path.join(d, x) for x in ("ref.fa", "snp.vcf", "reads.fq", "snp.txt"))
    synth.write_fasta(g, fa)
    synth.write_vcf(g, snps, vcf)
    synth.write_ucsc_txt(g, snps, txt, seed=12)
    haps = synth.donor_haplotypes(g, snps, seed=12)
    synth.simulate_reads(g, haps, 6000, 150, seed=12, sub_rate=0.01, lowq_prob=0.4).tofile(fq)
    return Dataset("ucscA", d, fa, vcf, fq, 6000, txt)


def make_big100(d: str) -> Dataset:
    """105 Mbp in two contigs with 400 repeat families (2..14 copies of 300..3300 bases, 0..2 mutations per copy: thousands of
    aux rows and POS_AMBIGUOUS entries whose column ORDER depends on how the reference's qsort treats equal k-mers), a planted
    16-base motif (big HI32 block) and 300 k SNPs: the size at which the index builder's byte-identity is pinned
    (tests/golden/big100.json holds the sha256 of the files the compiled reference wrote).  No reads."""
    fams = []
    for f in range(400):
        r = int(synth.rnd64(101, 50, f))
        fams.append((300 + r % 3000, 2 + (r >> 20) % 13, (r >> 30) % 3, f + 1))
    g = synth.make_genome([("chr1", 60_000_000), ("chr2", 45_000_000)], seed=101,
                          n_runs=[(0, 0, 10000), (0, 30_000_000, 1_500_000), (1, 20_000_000, 50_000), (1, 44_990_000, 10_000)],
                          repeats=fams, motifs=[(b"TTGACCGATAGGCATC", 150)])
    snps = synth.make_snps(g, 300_000, seed=101, cluster_frac=0.1)
    os.makedirs(d, exist_ok=True)
    fa, vcf, fq = (os.path.join(d, x) for x in ("ref.fa", "snp.vcf", "reads.fq"))
    synth.write_fasta(g, fa)
    synth.write_vcf(g, snps, vcf)
    open(fq, "wb").close()
    return Dataset("big100", d, fa, vcf, fq, 0)


MAKERS = {"s0": make_s0, "advA": make_adv_a, "advB": make_adv_b, "big100": make_big100, "ucscA": make_ucsc_a}
