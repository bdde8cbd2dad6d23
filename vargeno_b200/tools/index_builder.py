"""Byte-identical re-implementation of `vargeno index` (numpy) + index file I/O.

SURVEY.md section 8(f)-1 ("next" row): the on-disk index format stays exactly the reference's, so this
builder is validated byte-for-byte against files written by the compiled reference
(tests/test_index_builder.py, tests/golden/*.json hold the sha256 of reference-built files).
It exists because tests and bench.py need indexes on a fresh box in seconds, not because the
reference's offline path is being replaced.

Reference being restated (what, not how):
  ref dict   src/dictgen.c:12-51 (k-mer walk), :63-154 (collapse + write), :277-301
  snp dict   src/dictgen.c:561-794 (VCF walk), :156-275 (collapse + write)
  ref BF     src/generate_bf.cc:90-172     snp BF  src/generate_bf.cc:179-277 (incl. the discarded
             shift_kmer result at :257 -- every SNP inserts one value)
  chrlens    src/qv.cc:2344-2346           FASTA   src/fasta_parser.c:35-131, src/generate_bf.cc:18-76
  containers src/generate_bf.h:83-89 -> sdsl bit_vector: u64 bit count + u64 words
"""
from __future__ import annotations

import dataclasses
import re
import sys
from typing import List, Optional, Sequence, Tuple

import numpy as np

U64 = np.uint64
U32 = np.uint32

REF_BF_BITS = 1200000000 * 8       # src/generate_bf.h:201 REF_BF_BYTES
REF_LITE_BF_BITS = 2300000000 * 8  # src/generate_bf.h:202 (written, never read)
SNP_BF_BITS = 140000000 * 8        # src/generate_bf.h:203
POS_AMBIGUOUS = 0xFFFFFFFF         # src/vartype.h:33
AUX_COLS = 10                      # src/vartype.h:93

REF_REC = np.dtype([("kmer", "<u8"), ("pos", "<u4"), ("flag", "u1")])                      # 13 B
SNP_REC = np.dtype([("kmer", "<u8"), ("pos", "<u4"), ("snp", "u1"), ("flag", "u1"),
                    ("ref_freq", "u1"), ("alt_freq", "u1")])                                 # 16 B
SNP_AUX_COL = np.dtype([("pos", "<u4"), ("snp", "u1"), ("ref_freq", "u1"), ("alt_freq", "u1")])  # 7 B
SNP_AUX_REC = np.dtype([("kmer", "<u8"), ("cols", SNP_AUX_COL, (AUX_COLS,))])               # 78 B
assert REF_REC.itemsize == 13 and SNP_REC.itemsize == 16 and SNP_AUX_REC.itemsize == 78

_CODE = np.full(256, 7, dtype=np.uint8)   # BASE_X
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[_c + 32] = _i
_CODE[ord("N")] = 4
_CODE[ord("n")] = 4


@dataclasses.dataclass
class Index:
    """In-memory image of the five index files of one prefix."""
    ref: np.ndarray            # REF_REC[n]
    ref_aux: np.ndarray        # uint32[aux_n, 10]
    snp: np.ndarray            # SNP_REC[m]
    snp_aux: np.ndarray        # SNP_AUX_REC[aux_m]
    ref_bf_bits: int
    ref_bf: np.ndarray         # uint64 words; may be shorter than ceil(bits/64): missing words are 0
    snp_bf_bits: int
    snp_bf: np.ndarray         # uint64 words
    chr_names: List[str]
    chr_lens: List[int]
    ref_lite_bf: Optional[np.ndarray] = None  # only kept when asked for (2.3 GB)


# --------------------------------------------------------------------------------------
# hashes (src/generate_bf.h:112-142)
# --------------------------------------------------------------------------------------
def hash32(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=U32)
    with np.errstate(over="ignore"):
        x = ((x >> U32(16)) ^ x) * U32(0x45D9F3B)
        x = ((x >> U32(16)) ^ x) * U32(0x45D9F3B)
        x = (x >> U32(16)) ^ x
    return x


def hash40(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=U64)
    with np.errstate(over="ignore"):
        x = (x ^ (x >> U64(30))) * U64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> U64(27))) * U64(0x94D049BB133111EB)
        x = x ^ (x >> U64(31))
    return x


def _set_bits(words: np.ndarray, bit_idx: np.ndarray) -> None:
    if bit_idx.size == 0:
        return
    b = np.unique(np.asarray(bit_idx, dtype=U64))
    w = (b >> U64(6)).astype(np.int64)
    m = U64(1) << (b & U64(63))
    starts = np.flatnonzero(np.concatenate([[True], w[1:] != w[:-1]]))
    words[w[starts]] |= np.bitwise_or.reduceat(m, starts)


# --------------------------------------------------------------------------------------
# FASTA (two readers, as in the reference)
# --------------------------------------------------------------------------------------
def read_fasta_raw(path: str) -> Tuple[List[str], List[np.ndarray]]:
    """generate_bf.cc:18-76 view: name = whole header line after '>', raw characters, blank lines dropped."""
    data = open(path, "rb").read()
    names, seqs = [], []
    for block in data.split(b">")[1:]:
        nl = block.find(b"\n")
        if nl < 0:
            nl = len(block)
        names.append(block[:nl].decode("latin-1"))
        body = np.frombuffer(block[nl + 1:], dtype=np.uint8)
        seqs.append(body[body != 10].copy())
    return names, seqs


def normalise_fasta(names: Sequence[str], seqs: Sequence[np.ndarray]) -> Tuple[List[str], List[np.ndarray]]:
    """fasta_parser.c view: name cut at first whitespace or '|' (<= 64 chars), non-ACGT -> 'N', upper case."""
    out_n, out_s = [], []
    up = np.full(256, ord("N"), dtype=np.uint8)
    for c in b"ACGT":
        up[c] = c
        up[c + 32] = c
    for n, s in zip(names, seqs):
        m = re.match(r"[^\s|]{0,64}", n)
        out_n.append(m.group(0))
        out_s.append(up[s])
    return out_n, out_s


# --------------------------------------------------------------------------------------
# k-mers
# --------------------------------------------------------------------------------------
def contig_kmers(seq: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """All 32-mers of one contig: (kmer[u64] per start, valid[bool]); base b of a k-mer at bits 2b..2b+1
    (src/util.c:89-111).  valid = window free of N/n.  Characters outside ACGTNacgtn make the reference
    abort (util.c:103); rejected here."""
    code = _CODE[seq]
    if np.any(code == 7):
        raise ValueError("FASTA has characters outside ACGTN (reference encode_kmer would abort, src/util.c:103)")
    n = seq.size
    if n < 32:
        raise ValueError("contig shorter than 32 (reference asserts, src/dictgen.c:17)")
    k = (code & 3).astype(U64)
    for step in (1, 2, 4, 8, 16):
        k = k[:-step] | (k[step:] << U64(2 * step))
    isn = (code == 4).astype(np.int32)
    cs = np.concatenate([[0], np.cumsum(isn)])
    valid = (cs[32:] - cs[:-32]) == 0
    return k, valid


def _collapse(kmers_sorted: np.ndarray):
    """Group equal neighbours of a sorted k-mer array -> (group_start, group_size)."""
    n = kmers_sorted.size
    if n == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    gs = np.flatnonzero(np.concatenate([[True], kmers_sorted[1:] != kmers_sorted[:-1]]))
    gsz = np.diff(np.concatenate([gs, [n]]))
    return gs, gsz


def build_ref_dict(seqs: Sequence[np.ndarray], want_kmers: bool = False):
    """src/dictgen.c:277-301 + :63-154.  Returns (REF_REC[n], aux uint32[aux_n,10], per-contig kmers/valid)."""
    per_contig = []
    all_k, all_p = [], []
    start = 1  # 1-based position of the first base of the contig in the concatenation (dictgen.c:289)
    for s in seqs:
        k, valid = contig_kmers(s)
        per_contig.append((k, valid))
        idx = np.flatnonzero(valid)
        all_k.append(k[idx])
        all_p.append(((idx + start) & 0xFFFFFFFF).astype(U32))
        start += s.size
    k = np.concatenate(all_k)
    p = np.concatenate(all_p)
    order = np.argsort(k, kind="stable")  # glibc qsort = stable merge sort here
    k, p = k[order], p[order]
    gs, gsz = _collapse(k)
    rec = np.zeros(gs.size, dtype=REF_REC)
    rec["kmer"] = k[gs]
    small = (gsz >= 2) & (gsz <= AUX_COLS)
    big = gsz > AUX_COLS
    aux_idx = np.cumsum(small) - 1
    rec["pos"] = np.where(big, U32(POS_AMBIGUOUS), np.where(small, aux_idx.astype(U32), p[gs]))
    rec["flag"] = (gsz >= 2).astype(np.uint8)
    n_aux = int(small.sum())
    aux = np.zeros((n_aux, AUX_COLS), dtype=U32)
    sg, ssz = gs[small], gsz[small]
    for c in range(AUX_COLS):
        m = ssz > c
        aux[m, c] = p[sg[m] + c]
    return rec, aux, (per_contig if want_kmers else None)


@dataclasses.dataclass
class SnpLines:
    """The VCF records that survive the dictionary-side filters of src/dictgen.c:599-748, in file order."""
    contig: np.ndarray   # int32
    index0: np.ndarray   # int64 0-based in contig
    ref_code: np.ndarray  # uint8 0..3
    alt_code: np.ndarray  # uint8 0..3
    ref_freq: np.ndarray  # uint8
    alt_freq: np.ndarray  # uint8


_ATOF = re.compile(rb"^[ \t\n\v\f\r]*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?)")


def _atof(b: bytes) -> float:
    m = _ATOF.match(b)
    return float(m.group(0)) if m else 0.0


def freq_enc(f: float) -> int:
    """(uint8_t)(freq*0xff) with freq a C float (src/dictgen.c:741-742)."""
    return int(np.float32(f) * np.float32(255.0)) & 0xFF


def parse_vcf_for_dict(vcf_path: str, names: Sequence[str], seqs: Sequence[np.ndarray]) -> SnpLines:
    """Record-by-record restatement of the VCF walk of src/dictgen.c:599-748 (filters, CAF, fatal REF mismatch)."""
    ref_has_chr = names[0][:1] == "c"
    lens = [int(s.size) for s in seqs]
    name_to_idx = {}
    for i, n in enumerate(names):
        name_to_idx.setdefault(n, i)
    out = ([], [], [], [], [], [])
    has_freq, freq_index = True, -1
    isspace = b" \t\n\v\f\r"
    acgt = b"ACGT"
    with open(vcf_path, "rb") as f:
        for line in f:
            if line[:1] == b"#" or line[:1] == b"\n":
                continue
            if not line.endswith(b"\n"):
                line += b"\n"
            fld = line.split(b"\t")
            chrom = re.match(rb"[^ \t\n\v\f\r]*", fld[0]).group(0)
            if chrom[:1] != b"c" and ref_has_chr:
                chrom = (b"chr" + chrom)[:49]
            else:
                chrom = chrom[:49]
            ref_b = fld[3][:1].upper()
            rc = int(_CODE[ref_b[0]]) if ref_b else 7
            if rc == 7:
                continue
            nxt = fld[3][1:2]
            if nxt and nxt not in isspace:       # REF longer than one base (src/dictgen.c:644)
                continue
            alt_f = fld[4]
            nxt = alt_f[1:2]
            if nxt and nxt not in isspace:       # ALT longer than one base (src/dictgen.c:649)
                continue
            ci = name_to_idx.get(chrom.decode("latin-1"))
            if ci is None:
                sys.stderr.write("[Error] chromosome name %s in VCF file not found in reference\n" % chrom.decode())
                continue
            index = (int(_atof(fld[1])) - 1) & 0xFFFFFFFF
            seq = seqs[ci]
            if index >= lens[ci] or bytes([seq[index]]).upper() != ref_b:
                raise ValueError("Mismatch between reference sequence and SNP file at 0-based index %d in %s "
                                 "(reference exits, src/dictgen.c:666-672)" % (index, names[ci]))
            if index < 32 or index + 32 > lens[ci]:
                continue
            a2 = alt_f[:1].upper()
            if not ref_b or ref_b not in acgt or not a2 or a2 not in acgt:
                continue
            f1 = f2 = 0.5
            if has_freq:
                info = re.match(rb"[^ \t\n]*", fld[7]).group(0) if len(fld) > 7 else b""
                toks = []
                pos = 0
                while pos < len(info):
                    toks.append(info[pos:])
                    m = re.search(rb"[;=]", info[pos:])
                    if not m:
                        break
                    pos += m.end()
                for i, t in enumerate(toks):
                    if t.startswith(b"CAF"):
                        freq_index = i + 1
                if freq_index == -1:
                    has_freq = False
            if has_freq:
                if freq_index >= len(toks):
                    raise ValueError("record without CAF after one with CAF: undefined in the reference (src/dictgen.c:733-737)")
                p = toks[freq_index]
                f1 = _atof(p)
                comma = p.find(b",")
                if comma < 0:
                    raise ValueError("CAF without a comma: undefined in the reference (src/dictgen.c:736)")
                f2 = _atof(p[comma + 1:])
            if a2 == ref_b:
                continue
            win = seq[index - 32:index + 32]
            code = _CODE[win]
            if np.any(code[:32] == 7):
                raise ValueError("non-ACGTN in reference window")
            if np.any(code[:32] == 4):
                continue
            if np.any((win[33:] == ord("N")) | (win[33:] == ord("n"))):
                continue
            out[0].append(ci)
            out[1].append(index)
            out[2].append(rc)
            out[3].append(int(_CODE[a2[0]]))
            out[4].append(freq_enc(f1))
            out[5].append(freq_enc(f2))
    return SnpLines(np.array(out[0], np.int32), np.array(out[1], np.int64), np.array(out[2], np.uint8),
                    np.array(out[3], np.uint8), np.array(out[4], np.uint8), np.array(out[5], np.uint8))


def build_snp_dict(snps: SnpLines, seqs: Sequence[np.ndarray], per_contig_kmers) -> Tuple[np.ndarray, np.ndarray]:
    """src/dictgen.c:750-781 (32 alt-allele k-mers per SNP) + :156-275 (collapse + write)."""
    starts1 = np.concatenate([[1], 1 + np.cumsum([s.size for s in seqs])[:-1]]).astype(np.int64)
    n = snps.index0.size
    if n == 0:
        return np.zeros(0, SNP_REC), np.zeros(0, SNP_AUX_REC)
    i = np.arange(32, dtype=np.int64)[None, :]
    kstart = snps.index0[:, None] - 31 + i                      # k-mer i starts here (0-based in contig)
    refk = np.empty((n, 32), dtype=U64)
    for ci in np.unique(snps.contig):
        m = snps.contig == ci
        refk[m] = per_contig_kmers[int(ci)][0][kstart[m]]
    off = (31 - i).astype(U64)                                  # SNP offset inside k-mer i
    sh = off * U64(2)
    km = (refk & ~(U64(3) << sh)) | (snps.alt_code.astype(U64)[:, None] << sh)
    pos = ((starts1[snps.contig][:, None] + kstart) & 0xFFFFFFFF).astype(U32)
    info = ((off.astype(np.uint8) & 0x1F) << 3) | (snps.ref_code[:, None] & 7)
    km, pos, info = km.reshape(-1), pos.reshape(-1), np.broadcast_to(info, (n, 32)).reshape(-1)
    rf = np.repeat(snps.ref_freq, 32)
    af = np.repeat(snps.alt_freq, 32)
    order = np.argsort(km, kind="stable")
    km, pos, info, rf, af = km[order], pos[order], info[order], rf[order], af[order]
    gs, gsz = _collapse(km)
    rec = np.zeros(gs.size, dtype=SNP_REC)
    rec["kmer"] = km[gs]
    one = gsz == 1
    small = (gsz >= 2) & (gsz <= AUX_COLS)
    big = gsz > AUX_COLS
    aux_idx = np.cumsum(small) - 1
    rec["pos"] = np.where(big, U32(POS_AMBIGUOUS), np.where(small, aux_idx.astype(U32), pos[gs]))
    rec["snp"] = np.where(one, info[gs], 0)
    rec["flag"] = (~one).astype(np.uint8)
    rec["ref_freq"] = np.where(one, rf[gs], 0)
    rec["alt_freq"] = np.where(one, af[gs], 0)
    aux = np.zeros(int(small.sum()), dtype=SNP_AUX_REC)
    sg, ssz = gs[small], gsz[small]
    aux["kmer"] = km[sg]
    for c in range(AUX_COLS):
        m = ssz > c
        j = sg[m] + c
        aux["cols"]["pos"][m, c] = pos[j]
        aux["cols"]["snp"][m, c] = info[j]
        aux["cols"]["ref_freq"][m, c] = rf[j]
        aux["cols"]["alt_freq"][m, c] = af[j]
    return rec, aux


def build_ref_bf(per_contig_kmers, want_lite: bool = False):
    """src/generate_bf.cc:90-172.  hash32 % 9.6e9 is the identity (9.6e9 > 2^32): only the first 2^32 bits
    (2^26 words) can ever be set, so only those are materialised."""
    words = np.zeros(1 << 26, dtype=U64)
    lite = np.zeros((REF_LITE_BF_BITS + 63) // 64, dtype=U64) if want_lite else None
    for k, valid in per_contig_kmers:
        kv = k[valid]
        _set_bits(words, hash32((kv & U64(0xFFFFFFFF)).astype(U32)).astype(U64) % U64(REF_BF_BITS))
        if want_lite:
            _set_bits(lite, hash40(kv & U64(0xFFFFFFFFFF)) % U64(REF_LITE_BF_BITS))
    return words, lite


def build_snp_bf(vcf_path: str, raw_names: Sequence[str], raw_seqs: Sequence[np.ndarray]) -> np.ndarray:
    """src/generate_bf.cc:179-277, literally: the loop at :247-262 never stores shift_kmer's result, so each
    accepted record sets the single bit of LO40(32-mer ending just before the SNP).  The stale-`seq` behaviour
    for unknown contigs (:213-221) is kept too."""
    words = np.zeros((SNP_BF_BITS + 63) // 64, dtype=U64)
    vals = []
    pre, seq = "XO", np.zeros(0, np.uint8)
    with open(vcf_path, "rb") as f:
        for line in f.read().split(b"\n"):
            if not line or line[:1] == b"#":
                continue
            col = line.split(b"\t")
            chrom = col[0].decode("latin-1")
            if chrom[:1] != "c":
                chrom = "chr" + chrom
            pos = int(re.match(rb"\s*[+-]?\d+", col[1]).group(0)) - 1
            r, a = col[3], col[4]
            if len(r) > 1 or len(a) > 1:
                continue
            if chrom != pre:
                for nm, s in zip(raw_names, raw_seqs):
                    if nm == chrom:
                        seq = s
                        break
                pre = chrom
            if pos < 32 or pos + 32 > seq.size:
                continue
            if r[0] != seq[pos] or r == a:
                continue
            code = _CODE[seq[pos - 32:pos]]
            if np.any(code == 7):
                raise ValueError("non-ACGTN before SNP (reference aborts)")
            if np.any(code == 4):
                continue
            if a in (b"N", b"n"):
                continue
            if _CODE[a[0]] == 7:
                raise ValueError("ALT base %r makes the reference abort in shift_kmer (src/util.c:122)" % a)
            k = 0
            for b in range(32):
                k |= int(code[b] & 3) << (2 * b)
            vals.append(k & 0xFFFFFFFFFF)
    if vals:
        _set_bits(words, hash40(np.array(vals, dtype=U64)) % U64(SNP_BF_BITS))
    return words


def snp_bf_from_arrays(snps_contig, snps_index0, per_contig_kmers) -> np.ndarray:
    """Vectorised twin of build_snp_bf for records already known to pass its filters (synthetic sets)."""
    words = np.zeros((SNP_BF_BITS + 63) // 64, dtype=U64)
    vals = []
    for ci in np.unique(snps_contig):
        m = snps_contig == ci
        k, valid = per_contig_kmers[int(ci)]
        st = snps_index0[m] - 32
        st = st[valid[st]]
        vals.append(k[st] & U64(0xFFFFFFFFFF))
    if vals:
        _set_bits(words, hash40(np.concatenate(vals)) % U64(SNP_BF_BITS))
    return words


# --------------------------------------------------------------------------------------
# top level
# --------------------------------------------------------------------------------------
def build_index(fasta_path: str, vcf_path: str, want_lite: bool = False) -> Index:
    raw_names, raw_seqs = read_fasta_raw(fasta_path)
    names, seqs = normalise_fasta(raw_names, raw_seqs)
    ref, ref_aux, pck = build_ref_dict(seqs, want_kmers=True)
    lines = parse_vcf_for_dict(vcf_path, names, seqs)
    snp, snp_aux = build_snp_dict(lines, seqs, pck)
    # the BF side walks the raw FASTA; with bare upper-case ACGTN input the k-mers are the same ones
    same = all(np.array_equal(a, b) for a, b in zip(raw_seqs, seqs))
    pck_raw = pck if same else [contig_kmers(s) for s in raw_seqs]
    ref_bf, lite = build_ref_bf(pck_raw, want_lite)
    snp_bf = build_snp_bf(vcf_path, raw_names, raw_seqs)
    return Index(ref, ref_aux, snp, snp_aux, REF_BF_BITS, ref_bf, SNP_BF_BITS, snp_bf,
                 list(names), [int(s.size) for s in seqs], lite)


def _write_bv(path: str, bits: int, words: np.ndarray) -> None:
    nwords = (bits + 63) // 64
    with open(path, "wb") as f:
        f.write(np.array([bits], dtype="<u8").tobytes())
        w = np.ascontiguousarray(words[:nwords], dtype="<u8")
        w.tofile(f)
        rest = nwords - w.size
        if rest > 0:
            f.seek(rest * 8 - 1, 1)
            f.write(b"\0")


def write_index(ix: Index, prefix: str, write_lite: bool = False) -> None:
    with open(prefix + ".ref.dict", "wb") as f:
        f.write(np.array([ix.ref.size, ix.ref_aux.shape[0]], dtype="<u8").tobytes())
        ix.ref.tofile(f)
        np.ascontiguousarray(ix.ref_aux, dtype="<u4").tofile(f)
    with open(prefix + ".snp.dict", "wb") as f:
        f.write(np.array([ix.snp.size, ix.snp_aux.size], dtype="<u8").tobytes())
        ix.snp.tofile(f)
        ix.snp_aux.tofile(f)
    _write_bv(prefix + ".ref.bf", ix.ref_bf_bits, ix.ref_bf)
    _write_bv(prefix + ".snp.bf", ix.snp_bf_bits, ix.snp_bf)
    if write_lite:
        assert ix.ref_lite_bf is not None
        _write_bv(prefix + ".ref.bf.lite.bf", REF_LITE_BF_BITS, ix.ref_lite_bf)
    with open(prefix + ".chrlens", "w") as f:
        for n, l in zip(ix.chr_names, ix.chr_lens):
            f.write("%s %d\n" % (n, l))


def _read_bv(path: str, max_words: Optional[int] = None) -> Tuple[int, np.ndarray]:
    bits = int(np.fromfile(path, dtype="<u8", count=1)[0])
    nwords = (bits + 63) // 64
    if max_words is not None:
        nwords = min(nwords, max_words)
    return bits, np.fromfile(path, dtype="<u8", count=nwords, offset=8)


def load_index(prefix: str) -> Index:
    """Reader for the files `vargeno index` writes (layout: SURVEY.md 3.3 table; src/qv.cc:519-695)."""
    hdr = np.fromfile(prefix + ".ref.dict", dtype="<u8", count=2)
    n, an = int(hdr[0]), int(hdr[1])
    ref = np.fromfile(prefix + ".ref.dict", dtype=REF_REC, count=n, offset=16)
    ref_aux = np.fromfile(prefix + ".ref.dict", dtype="<u4", count=an * AUX_COLS, offset=16 + 13 * n).reshape(an, AUX_COLS)
    hdr = np.fromfile(prefix + ".snp.dict", dtype="<u8", count=2)
    m, am = int(hdr[0]), int(hdr[1])
    snp = np.fromfile(prefix + ".snp.dict", dtype=SNP_REC, count=m, offset=16)
    snp_aux = np.fromfile(prefix + ".snp.dict", dtype=SNP_AUX_REC, count=am, offset=16 + 16 * m)
    rbits, rbf = _read_bv(prefix + ".ref.bf", max_words=1 << 26)  # hash32 cannot address beyond 2^32 bits
    sbits, sbf = _read_bv(prefix + ".snp.bf")
    names, lens = [], []
    for line in open(prefix + ".chrlens"):
        if line.strip():
            a, b = line.split()[:2]
            names.append(a)
            lens.append(int(b))
    return Index(ref, ref_aux, snp, snp_aux, rbits, rbf, sbits, sbf, names, lens)


# --------------------------------------------------------------------------------------
# array fast path for synthetic sets (bench.py): no FASTA / VCF text round trip
# --------------------------------------------------------------------------------------
def snplines_from_arrays(seqs: Sequence[np.ndarray], contig, pos0, ref_ascii, alt_ascii, freq1, freq2) -> SnpLines:
    """Vectorised twin of parse_vcf_for_dict for well-formed bi-allelic ACGT records that all carry CAF
    (what tools/synth.write_vcf emits without extra lines).  freq1 / freq2 are the two CAF numbers as printed."""
    contig = np.asarray(contig, np.int32)
    pos0 = np.asarray(pos0, np.int64)
    keep = np.zeros(pos0.size, bool)
    for ci in np.unique(contig):
        m = np.flatnonzero(contig == ci)
        s = seqs[int(ci)]
        p = pos0[m]
        ok = (p >= 32) & (p + 32 <= s.size)
        if np.any(s[p[ok]] != np.asarray(ref_ascii)[m][ok]):
            raise ValueError("REF does not match the reference sequence (reference exits, src/dictgen.c:666-672)")
        isn = np.concatenate([[0], np.cumsum((_CODE[s] == 4).astype(np.int64))])
        q = np.where(ok, p, 32)
        nfree = (isn[q + 32] - isn[q - 32]) == 0            # bases index-32 .. index+31 (src/dictgen.c:756-765)
        keep[m] = ok & nfree
    keep &= np.asarray(ref_ascii) != np.asarray(alt_ascii)
    f1 = (np.asarray(freq1, np.float64).astype(np.float32) * np.float32(255.0)).astype(np.int64) & 0xFF
    f2 = (np.asarray(freq2, np.float64).astype(np.float32) * np.float32(255.0)).astype(np.int64) & 0xFF
    return SnpLines(contig[keep], pos0[keep], _CODE[np.asarray(ref_ascii)][keep], _CODE[np.asarray(alt_ascii)][keep],
                    f1[keep].astype(np.uint8), f2[keep].astype(np.uint8))


def build_index_from_arrays(names: Sequence[str], seqs: Sequence[np.ndarray], contig, pos0, ref_ascii, alt_ascii,
                            freq1, freq2) -> Index:
    """Same Index as build_index(write_fasta(...), write_vcf(...)) for upper-case ACGTN contigs with bare names."""
    ref, ref_aux, pck = build_ref_dict(seqs, want_kmers=True)
    lines = snplines_from_arrays(seqs, contig, pos0, ref_ascii, alt_ascii, freq1, freq2)
    snp, snp_aux = build_snp_dict(lines, seqs, pck)
    ref_bf, _ = build_ref_bf(pck, False)
    # BF-side filter (src/generate_bf.cc:224-241): position window, REF matches, REF != ALT, 32 bases before N-free
    contig = np.asarray(contig, np.int32)
    pos0 = np.asarray(pos0, np.int64)
    ok = np.zeros(pos0.size, bool)
    for ci in np.unique(contig):
        m = contig == ci
        ok[m] = (pos0[m] >= 32) & (pos0[m] + 32 <= seqs[int(ci)].size)
    ok &= np.asarray(ref_ascii) != np.asarray(alt_ascii)
    snp_bf = snp_bf_from_arrays(contig[ok], pos0[ok], pck)
    return Index(ref, ref_aux, snp, snp_aux, REF_BF_BITS, ref_bf, SNP_BF_BITS, snp_bf, list(names), [int(s.size) for s in seqs])


def bf_line_positions(vcf_path: str, raw_names: Sequence[str], raw_seqs: Sequence[np.ndarray]):
    """(contig index, 0-based position) of the VCF lines that reach the insert loop of constructBfFromVcf
    (src/generate_bf.cc:203-246) apart from the N test of the preceding 32-mer, which the GPU builder does itself.
    Same walk as build_snp_bf, including the stale-sequence behaviour for unknown contigs."""
    out_c, out_p = [], []
    pre, ci, seq = "XO", -1, np.zeros(0, np.uint8)
    with open(vcf_path, "rb") as f:
        for line in f.read().split(b"\n"):
            if not line or line[:1] == b"#":
                continue
            col = line.split(b"\t")
            chrom = col[0].decode("latin-1")
            if chrom[:1] != "c":
                chrom = "chr" + chrom
            pos = int(re.match(rb"\s*[+-]?\d+", col[1]).group(0)) - 1
            r, a = col[3], col[4]
            if len(r) > 1 or len(a) > 1:
                continue
            if chrom != pre:
                for k, (nm, s) in enumerate(zip(raw_names, raw_seqs)):
                    if nm == chrom:
                        seq, ci = s, k
                        break
                pre = chrom
            if pos < 32 or pos + 32 > seq.size:
                continue
            if r[0] != seq[pos] or r == a or a in (b"N", b"n"):
                continue
            if _CODE[a[0]] == 7:
                raise ValueError("ALT base %r makes the reference abort in shift_kmer (src/util.c:122)" % a)
            out_c.append(ci)
            out_p.append(pos)
    return np.array(out_c, np.int64), np.array(out_p, np.int64)
