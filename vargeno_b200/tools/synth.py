"""Deterministic synthetic inputs for the `vargeno geno` hot path (FASTA / VCF / FASTQ).

Tooling, not product: used by tests/ and bench.py to make inputs of the shapes SURVEY.md
section 8(d) names (S0 smoke, S1 chr22-shaped, adversarial sets A/B).  Everything is driven
by a counter-based RNG (splitmix64 finaliser over (seed, stream, a, b)) instead of
numpy.random, so the byte streams do not depend on the numpy version and can be reproduced
on the device.

Input contract the generators respect (reference behaviour):
  * FASTA: bare ">chrN" headers, alphabet ACGTN, fixed line width      (src/fasta_parser.c:35,
    src/generate_bf.cc:18 read the same file with two different parsers)
  * VCF:   every record carries CAF=<ref>,<alt>                          (src/dictgen.c:709-738)
  * FASTQ: 4 lines per record, file ends with exactly one '\n'           (src/qv.cc:760-763)
"""
from __future__ import annotations

import dataclasses
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

U64 = np.uint64
_MASK = (1 << 64) - 1

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i  # lower case


def rnd64(seed: int, stream: int, a, b=0) -> np.ndarray:
    """Counter-based 64-bit random value(s): splitmix64 finaliser of a mix of the four keys.

    `a` and `b` may be scalars or uint64 arrays (broadcast).  Pure integer arithmetic, wraps
    mod 2^64 -- the CUDA twin (csrc/synth_reads.cuh) evaluates the same expression.
    """
    with np.errstate(over="ignore"):
        a = np.asarray(a, dtype=U64)
        b = np.asarray(b, dtype=U64)
        x = U64((seed + 0x9E3779B97F4A7C15 * (stream + 1)) & _MASK)
        x = x ^ (a * U64(0xBF58476D1CE4E5B9))
        x = x + (b * U64(0x94D049BB133111EB))
        x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
        x = x ^ (x >> U64(31))
    return x


# --------------------------------------------------------------------------------------
# genome
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class Genome:
    names: List[str]
    seqs: List[np.ndarray]  # uint8 ASCII, upper case ACGTN
    repeat_copies: List[Tuple[int, int, int, int, int]] = dataclasses.field(default_factory=list)  # (family, copy, contig, offset, len)
    motif_sites: List[Tuple[int, int, int]] = dataclasses.field(default_factory=list)              # (motif, contig, offset)

    @property
    def lengths(self) -> List[int]:
        return [int(s.size) for s in self.seqs]

    @property
    def starts(self) -> np.ndarray:
        """0-based offset of each contig in the concatenation (dict positions are these + 1)."""
        return np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)

    def concat(self) -> np.ndarray:
        return np.concatenate(self.seqs) if len(self.seqs) > 1 else self.seqs[0]


def make_genome(
    contigs: Sequence[Tuple[str, int]],
    seed: int,
    n_runs: Sequence[Tuple[int, int, int]] = (),
    repeats: Sequence[Tuple[int, int, int, int]] = (),
    motifs: Sequence[Tuple[bytes, int]] = (),
) -> Genome:
    """Random ACGT contigs.

    n_runs:  (contig_idx, start, length) stretches overwritten with 'N'
    repeats: (family_len, copies, mutations_per_copy, family_id) -- a random source segment is
             copied `copies` times to random places (any contig), each copy with that many
             point mutations; exercises aux rows (2..10 copies) and POS_AMBIGUOUS (>10).
    motifs:  (motif_bytes, times) -- the motif is planted `times` times at random places;
             a 16-mer planted >=100 times makes a ref HI32 block >= BLOCK_SIZE_THRESHOLD.
    """
    seqs = []
    for ci, (_, length) in enumerate(contigs):
        r = rnd64(seed, 1, ci, np.arange(length, dtype=U64))
        seqs.append(BASES[(r >> U64(33)).astype(np.int64) & 3].copy())
    total = [s.size for s in seqs]
    rep_meta, mot_meta = [], []

    def place(stream, key, seg_len):
        r = int(rnd64(seed, stream, key))
        ci = r % len(seqs)
        off = (r >> 20) % max(1, total[ci] - seg_len)
        return ci, off

    for fam_len, copies, muts, fam_id in repeats:
        ci, off = place(2, fam_id, fam_len)
        src = seqs[ci][off:off + fam_len].copy()
        for c in range(copies):
            cj, o2 = place(3, fam_id * 1000 + c, fam_len)
            seg = src.copy()
            for m in range(muts):
                r = int(rnd64(seed, 4, fam_id * 1000 + c, m))
                p = r % fam_len
                seg[p] = BASES[(int(_CODE[seg[p]]) + 1 + ((r >> 40) % 3)) & 3]
            seqs[cj][o2:o2 + fam_len] = seg
            rep_meta.append((fam_id, c, cj, o2, fam_len))
    for mi, (motif, times) in enumerate(motifs):
        mot = np.frombuffer(motif, dtype=np.uint8)
        for t in range(times):
            cj, o2 = place(5, mi * 100000 + t, mot.size)
            seqs[cj][o2:o2 + mot.size] = mot
            mot_meta.append((mi, cj, o2))
    for ci, start, length in n_runs:
        seqs[ci][start:start + length] = ord("N")
    return Genome([n for n, _ in contigs], seqs, rep_meta, mot_meta)


def write_fasta(genome: Genome, path: str, width: int = 60) -> None:
    with open(path, "wb") as f:
        for name, seq in zip(genome.names, genome.seqs):
            f.write(b">" + name.encode() + b"\n")
            n = seq.size
            full = (n // width) * width
            if full:
                body = np.empty((full // width, width + 1), dtype=np.uint8)
                body[:, :width] = seq[:full].reshape(-1, width)
                body[:, width] = 10
                f.write(body.tobytes())
            if n > full:
                f.write(seq[full:].tobytes() + b"\n")


# --------------------------------------------------------------------------------------
# SNP list (VCF) + donor
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class SnpSet:
    contig: np.ndarray   # int32 contig index
    pos0: np.ndarray     # int64 0-based position inside the contig
    ref: np.ndarray      # uint8 ASCII
    alt: np.ndarray      # uint8 ASCII
    caf_ref: np.ndarray  # float64 as printed
    gt: np.ndarray       # uint8 donor genotype 0: 0/0, 1: 0/1, 2: 1/1


CAF_CHOICES = np.array([0.01, 0.1, 0.3, 0.5])


def make_snps(genome: Genome, n_snps: int, seed: int, cluster_frac: float = 0.0) -> SnpSet:
    """Bi-allelic SNPs at distinct positions >= 32 bases from contig ends, not on an N.

    cluster_frac: that fraction of the SNPs is placed 1..25 bases after another SNP
    (clustered sites: several SNPs inside one 32-mer).
    """
    lens = np.array(genome.lengths, dtype=np.int64)
    starts = genome.starts
    total = int(lens.sum())
    cat = genome.concat()
    n_draw = int(n_snps * 1.3) + 64
    g = (rnd64(seed, 10, np.arange(n_draw, dtype=U64)) % U64(total)).astype(np.int64)
    n_cl = int(n_snps * cluster_frac)
    if n_cl:
        base = g[:n_cl]
        delta = (rnd64(seed, 11, np.arange(n_cl, dtype=U64)) % U64(25)).astype(np.int64) + 1
        g = np.concatenate([g, base + delta])
    g = np.unique(g)
    g = g[g < total]
    ci = np.searchsorted(starts, g, side="right") - 1
    p0 = g - starts[ci]
    ok = (p0 >= 32) & (p0 + 32 <= lens[ci]) & (cat[g] != ord("N"))
    g, ci, p0 = g[ok], ci[ok], p0[ok]
    # deterministic thinning to n_snps, keep sorted order
    if g.size > n_snps:
        key = rnd64(seed, 12, g.astype(U64))
        keep = np.sort(np.argsort(key, kind="stable")[:n_snps])
        g, ci, p0 = g[keep], ci[keep], p0[keep]
    ref = cat[g]
    r = rnd64(seed, 13, g.astype(U64))
    alt = BASES[(_CODE[ref].astype(np.int64) + 1 + ((r >> U64(20)) % U64(3)).astype(np.int64)) & 3]
    caf = CAF_CHOICES[((r >> U64(30)) % U64(4)).astype(np.int64)]
    gt = ((r >> U64(40)) % U64(3)).astype(np.uint8)
    return SnpSet(ci.astype(np.int32), p0, ref, alt, caf, gt)


def write_vcf(
    genome: Genome,
    snps: SnpSet,
    path: str,
    chrom_prefix: bool = True,
    extra_lines: Optional[Dict[int, List[str]]] = None,
    sample_columns: bool = False,
    declare_gt: bool = False,
) -> None:
    """8-column VCF (or 10+ columns with sample_columns) with CAF in INFO.

    extra_lines: {snp_index: [raw record lines]} inserted right after that SNP's own line
    (second-allele lines, indel lines at a SNP POS ...).
    """
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.0\n##source=vargeno_b200.tools.synth\n")
        f.write('##INFO=<ID=CAF,Number=.,Type=String,Description="allele frequencies">\n')
        if declare_gt:
            f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        hdr = "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"
        if sample_columns:
            hdr += "\tFORMAT\tS1\tS2"
        f.write(hdr + "\n")
        out = []
        for i in range(snps.pos0.size):
            name = genome.names[int(snps.contig[i])]
            if not chrom_prefix and name.startswith("chr"):
                name = name[3:]
            cr = float(snps.caf_ref[i])
            line = "%s\t%d\trs%d\t%s\t%s\t.\t.\tRS=%d;CAF=%s,%s;COMMON=1" % (
                name, int(snps.pos0[i]) + 1, i + 1, chr(snps.ref[i]), chr(snps.alt[i]), i + 1,
                repr(round(1.0 - cr, 6)), repr(cr))
            if sample_columns:
                line += "\tDP:GT\t7:./.\tDP:GT\t9:0/1" if declare_gt else "\tDP\t7\tDP\t9"
            out.append(line)
            if extra_lines and i in extra_lines:
                out.extend(extra_lines[i])
        f.write("\n".join(out) + "\n")


def write_ucsc_txt(genome: Genome, snps: SnpSet, path: str, seed: int = 0, adversarial: bool = True) -> None:
    """UCSC snp141Common-style table dump (26 tab-separated columns; the legacy SNP input of the reference:
    src/dictgen.c:350-540 make_snp_dict, src/generate_bf.cc:439-592 constructBfFromUcsc).  Columns the reference reads:
    1 chrom, 2 chromStart (0-based), 6 strand, 7 refNCBI, 8 refUCSC, 9 observed, 11 class, 21 alleleFreqCount, 22 alleles,
    24 alleleFreqs.  About a third of the records are on the '-' strand (alleles complemented, as UCSC prints them), the
    allele order is shuffled (the frequency swap of :469-473), and with `adversarial` some records are of the kinds the
    filters drop: not "single", three alleles, multi-base reference, refNCBI != refUCSC, unknown contig, neither allele equal
    to the reference base, comment lines."""
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    out = ["#bin\tchrom\tchromStart\tchromEnd\tname\tscore\tstrand\trefNCBI\trefUCSC\tobserved\tmolType\tclass\tvalid\tavHet\tavHetSE\tfunc\t"
           "locType\tweight\texceptions\tsubmitterCount\tsubmitters\talleleFreqCount\talleles\talleleNs\talleleFreqs\tbitfields"]

    def row(chrom, start, strand, ref1, ref2, observed, klass, count, alleles, freqs, name):
        return "\t".join(["585", chrom, str(start), str(start + 1), name, "0", strand, ref1, ref2, observed, "genomic", klass, "by-frequency",
                          "0.3", "0.1", "intron", "exact", "1", "", "2", "A,B,", str(count), alleles, "10.0,20.0,", freqs, "maf-5-some-pop"])

    for i in range(snps.pos0.size):
        r = int(rnd64(seed, 70, i))
        name = genome.names[int(snps.contig[i])]
        ref, alt = chr(snps.ref[i]), chr(snps.alt[i])
        cr = float(snps.caf_ref[i])
        f_ref, f_alt = repr(round(1.0 - cr, 6)), repr(cr)
        neg = (r % 3) == 0
        a, b = (comp[ref], comp[alt]) if neg else (ref, alt)
        fa, fb = f_ref, f_alt
        if (r >> 8) & 1:
            a, b, fa, fb = b, a, fb, fa
        observed = "/".join(sorted([a, b]))
        start = int(snps.pos0[i])
        out.append(row(name, start, "-" if neg else "+", ref, ref, observed, "single", 2, a + "," + b + ",", fa + "," + fb + ",", "rs%d" % (i + 1)))
        if adversarial and i % 23 == 0:
            kind = (i // 23) % 7
            nxt = chr(genome.seqs[int(snps.contig[i])][start + 1]) if start + 1 < genome.seqs[int(snps.contig[i])].size else "A"
            others = [x for x in "ACGT" if x not in (ref, alt)]
            if kind == 0:
                out.append(row(name, start, "+", ref, ref, "-/" + ref, "deletion", 2, "-," + ref + ",", "0.5,0.5,", "rsD%d" % i))
            elif kind == 1:
                out.append(row(name, start, "+", ref, ref, "/".join(sorted([ref, alt, others[0]])), "single", 3,
                               ",".join([ref, alt, others[0]]) + ",", "0.5,0.25,0.25,", "rsT%d" % i))
            elif kind == 2 and nxt in "ACGT":
                out.append(row(name, start, "+", ref + nxt, ref + nxt, ref + nxt + "/" + ref, "in-del", 2, ref + nxt + "," + ref + ",", "0.5,0.5,", "rsM%d" % i))
            elif kind == 3:
                out.append(row(name, start, "+", ref, others[0], observed, "single", 2, a + "," + b + ",", fa + "," + fb + ",", "rsR%d" % i))
            elif kind == 4:
                out.append(row("chrUn_gl000%d" % (i % 9), start, "+", ref, ref, observed, "single", 2, a + "," + b + ",", fa + "," + fb + ",", "rsU%d" % i))
            elif kind == 5:
                out.append(row(name, start, "+", ref, ref, "/".join(sorted(others)), "single", 2, others[0] + "," + others[1] + ",", "0.5,0.5,", "rsX%d" % i))
            else:
                out.append("# a comment line in the middle of the table")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")


def donor_haplotypes(genome: Genome, snps: SnpSet, seed: int) -> Tuple[np.ndarray, np.ndarray]:
    """Two concatenated haplotype sequences of the donor (alt applied per genotype)."""
    cat = genome.concat()
    g = genome.starts[snps.contig] + snps.pos0
    h0, h1 = cat.copy(), cat.copy()
    hom = snps.gt == 2
    het = snps.gt == 1
    which = (rnd64(seed, 20, g.astype(U64)) >> U64(17)) & U64(1)
    h0[g[hom]] = snps.alt[hom]
    h1[g[hom]] = snps.alt[hom]
    m0 = het & (which == 0)
    m1 = het & (which == 1)
    h0[g[m0]] = snps.alt[m0]
    h1[g[m1]] = snps.alt[m1]
    return h0, h1


# --------------------------------------------------------------------------------------
# reads
# --------------------------------------------------------------------------------------
_COMP = np.zeros(256, dtype=np.uint8)
for _a, _b in zip(b"ACGTNacgtn", b"TGCANtgcan"):
    _COMP[_a] = _b

LOWQ_THRESHOLD = ord("8")  # src/vartype.h:17 QUALITY_SCORE


def simulate_reads(
    genome: Genome,
    haps: Tuple[np.ndarray, np.ndarray],
    n_reads: int,
    read_len: int,
    seed: int,
    sub_rate: float = 0.005,
    lowq_prob: float = 0.25,
    lowq_chars: int = 4,
    first_id: int = 0,
    id_width: int = 9,
    lower_rate: float = 0.0,
    n_rate: float = 0.0,
    forced_starts: Optional[np.ndarray] = None,
    forced_rev: Optional[np.ndarray] = None,
    qual_alphabet: Optional[bytes] = None,
) -> np.ndarray:
    """FASTQ text (uint8 array) of `n_reads` fixed-length records.

    Record r: "@r<id zero-padded>\\n<seq>\\n+\\n<qual>\\n"; uniform start, 50 % reverse strand,
    uniform substitutions at `sub_rate`, haplotype chosen per read.  Quality characters: each of
    the first `lowq_chars` is below '8' with probability `lowq_prob` (that is the character the
    reference gates the neighbour search of k-mer i on: src/qv.cc:836,943), all others are >= '8'.
    lower_rate / n_rate: per-base probability of lower-casing / replacing by 'N' (adversarial sets).
    forced_starts / forced_rev: per-read global 0-based start and strand instead of the random ones
    (targeted reads: put a planted motif on bases 16..31 of a k-mer).
    qual_alphabet: when given, every quality character is drawn uniformly from it instead.
    """
    lens = np.array(genome.lengths, dtype=np.int64)
    starts = genome.starts
    total = int(lens.sum())
    L = read_len
    assert all(l >= L for l in lens)
    rid = np.arange(first_id, first_id + n_reads, dtype=U64)
    r0 = rnd64(seed, 30, rid)
    s = (r0 % U64(total - L + 1)).astype(np.int64)
    ci = np.searchsorted(starts, s, side="right") - 1
    if forced_starts is not None:
        s = np.asarray(forced_starts, dtype=np.int64)
        ci = np.searchsorted(starts, s, side="right") - 1
    s = np.minimum(s, starts[ci] + lens[ci] - L)
    r1 = rnd64(seed, 31, rid)
    hap = ((r1 >> U64(7)) & U64(1)).astype(bool)
    rev = ((r1 >> U64(9)) & U64(1)).astype(bool)
    if forced_rev is not None:
        rev = np.asarray(forced_rev, dtype=bool)
    idx = s[:, None] + np.arange(L, dtype=np.int64)[None, :]
    seq = np.where(hap[:, None], haps[1][idx], haps[0][idx])
    # substitutions (in genome orientation, before strand flip)
    col = np.arange(L, dtype=U64)[None, :]
    rb = rnd64(seed, 32, rid[:, None], col)
    thr = U64(int(sub_rate * (1 << 32)))
    is_sub = ((rb & U64(0xFFFFFFFF)) < thr) & (seq != ord("N"))
    newb = BASES[(_CODE[seq].astype(np.int64) + 1 + ((rb >> U64(40)) % U64(3)).astype(np.int64)) & 3]
    seq = np.where(is_sub, newb, seq)
    seq = np.where(rev[:, None], _COMP[seq[:, ::-1]], seq)
    if n_rate > 0 or lower_rate > 0:
        rc = rnd64(seed, 33, rid[:, None], col)
        if n_rate > 0:
            seq = np.where((rc & U64(0xFFFFFFFF)) < U64(int(n_rate * (1 << 32))), np.uint8(ord("N")), seq)
        if lower_rate > 0:
            low = ((rc >> U64(32)) < U64(int(lower_rate * (1 << 32))))
            seq = np.where(low, seq | np.uint8(32), seq)
    rq = rnd64(seed, 34, rid[:, None], col)
    hi_q = (np.uint8(ord("8")) + ((rq >> U64(8)) % U64(19)).astype(np.uint8))          # '8'..'J'
    lo_q = (np.uint8(ord("#")) + ((rq >> U64(16)) % U64(21)).astype(np.uint8))         # '#'..'7'
    is_low = ((rq >> U64(32)) < U64(int(lowq_prob * (1 << 32)))) & (np.arange(L)[None, :] < lowq_chars)
    qual = np.where(is_low, lo_q, hi_q)
    if qual_alphabet is not None:
        qa = np.frombuffer(qual_alphabet, dtype=np.uint8)
        qual = qa[((rq >> U64(24)) % U64(qa.size)).astype(np.int64)]

    rec_len = 2 + id_width + 1 + L + 3 + L + 1
    out = np.empty((n_reads, rec_len), dtype=np.uint8)
    out[:, 0] = ord("@")
    out[:, 1] = ord("r")
    ids = np.arange(first_id, first_id + n_reads, dtype=np.int64)
    for d in range(id_width):
        out[:, 2 + id_width - 1 - d] = ord("0") + (ids // (10 ** d)) % 10
    o = 2 + id_width
    out[:, o] = 10
    out[:, o + 1:o + 1 + L] = seq
    o += 1 + L
    out[:, o] = 10
    out[:, o + 1] = ord("+")
    out[:, o + 2] = 10
    out[:, o + 3:o + 3 + L] = qual
    out[:, o + 3 + L] = 10
    return out.reshape(-1)


def fastq_records(text: bytes) -> List[Tuple[bytes, bytes, bytes, bytes]]:
    lines = text.split(b"\n")
    assert lines[-1] == b"" and (len(lines) - 1) % 4 == 0
    return [tuple(lines[i:i + 4]) for i in range(0, len(lines) - 1, 4)]
