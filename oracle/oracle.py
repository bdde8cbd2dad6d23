"""ctypes front-end of the CPU oracle (oracle/liboracle.so) + the oracle's VCF writer.

TEST INFRASTRUCTURE ONLY -- see oracle/vg_oracle.h.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline leg; never by vargeno_b200/.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from typing import Dict, List, Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "vargeno")
REF_INSTR_BIN = os.path.join(HERE, "_ref", "vargeno_instr")

READ_RESULT = np.dtype([("flags", "<u4"), ("target", "<u4"), ("freq", "<u2"), ("n_ref", "<u2"),
                        ("n_snp", "<u2"), ("passes", "<u2"), ("ctx_hash", "<u8")])
SITE = np.dtype([("pos", "<u4"), ("ref", "u1"), ("alt", "u1"), ("ref_cnt", "u1"), ("alt_cnt", "u1"),
                 ("ref_freq", "u1"), ("alt_freq", "u1"), ("pad0", "u1"), ("pad1", "u1")])
STATS_FIELDS = ["reads", "skipped_n", "passes", "placed", "exact_lookups", "nbr_query_lookups",
                "nbr_scan_reads", "bf_probes", "lowq_kmers", "events", "pileup_incr", "big_kmers"]
assert READ_RESULT.itemsize == 24 and SITE.itemsize == 12

F_SKIPPED, F_REVCOMPL, F_PROCESS, F_AMBIGUOUS, F_HASBEST = 1, 2, 4, 8, 16
GT_TEXT = {1: "0/0", 3: "0/1", 2: "1/1"}   # GTYPE_REF / GTYPE_HET / GTYPE_ALT (src/vartype.h:29-31, qv.cc:1678-1680)


def build(force: bool = False) -> str:
    """Compile liboracle.so (and nothing else) with the committed Makefile."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("vg_oracle.c", "vg_oracle.h")):
        subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.vgo_index_create.restype = C.c_void_p
        L.vgo_index_create.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64,
                                       C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                       C.c_void_p, C.c_uint64, C.c_uint64]
        L.vgo_index_free.argtypes = [C.c_void_p]
        L.vgo_lookup.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.vgo_blocks.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 4
        L.vgo_bf_check.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        L.vgo_process_fastq.restype = C.c_int64
        L.vgo_process_fastq.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_char_p]
        L.vgo_reset_pileup.argtypes = [C.c_void_p]
        L.vgo_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.vgo_get_sites.restype = C.c_uint64
        L.vgo_get_sites.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.vgo_add_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.vgo_call.argtypes = [C.c_int, C.c_int, C.c_uint8, C.c_uint8, C.POINTER(C.c_double)]
        L.vgo_gq.argtypes = [C.c_double]
        L.vgo_tables.argtypes = [C.c_void_p, C.c_void_p]
        L.vgo_ctx_digest.restype = C.c_uint64
        L.vgo_ctx_digest.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32]
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Oracle:
    """CPU oracle over an in-memory index image (vargeno_b200.tools.index_builder.Index)."""

    def __init__(self, ix):
        self.ix = ix
        L = lib()
        # keep contiguous byte images alive for the lifetime of the handle (the BF words are borrowed)
        self._ref = np.ascontiguousarray(ix.ref).view(np.uint8)
        self._ref_aux = np.ascontiguousarray(ix.ref_aux, dtype="<u4")
        self._snp = np.ascontiguousarray(ix.snp).view(np.uint8)
        self._snp_aux = np.ascontiguousarray(ix.snp_aux).view(np.uint8)
        self._rbf = np.ascontiguousarray(ix.ref_bf, dtype="<u8")
        self._sbf = np.ascontiguousarray(ix.snp_bf, dtype="<u8")
        self.h = L.vgo_index_create(_ptr(self._ref), ix.ref.size, _ptr(self._ref_aux), ix.ref_aux.shape[0],
                                    _ptr(self._snp), ix.snp.size, _ptr(self._snp_aux), ix.snp_aux.size,
                                    _ptr(self._rbf), ix.ref_bf_bits, self._rbf.size,
                                    _ptr(self._sbf), ix.snp_bf_bits, self._sbf.size)

    def close(self):
        if self.h:
            lib().vgo_index_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def lookup(self, which: int, kmer: int) -> Optional[Tuple[int, int, int]]:
        pos, flag, info = C.c_uint32(), C.c_uint8(), C.c_uint8()
        if lib().vgo_lookup(self.h, which, kmer, C.byref(pos), C.byref(flag), C.byref(info)):
            return pos.value, flag.value, info.value
        return None

    def blocks(self, kmer: int):
        v = [C.c_uint32() for _ in range(4)]
        lib().vgo_blocks(self.h, kmer, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)

    def bf_check(self, which: int, value: int) -> int:
        return lib().vgo_bf_check(self.h, which, value)

    def process_fastq(self, text, trace_path: Optional[str] = None, want_results: bool = True) -> np.ndarray:
        buf = np.frombuffer(text, dtype=np.uint8) if not isinstance(text, np.ndarray) else text
        cap = int(np.count_nonzero(buf == 10)) // 4 + 2 if want_results else 0
        res = np.zeros(cap, dtype=READ_RESULT)
        n = lib().vgo_process_fastq(self.h, _ptr(buf), buf.size, _ptr(res) if want_results else None, cap,
                                    trace_path.encode() if trace_path else None)
        if n < 0:
            raise ValueError("oracle rejected the FASTQ text: code %d" % n)
        return res[:n]

    def reset(self):
        lib().vgo_reset_pileup(self.h)

    def stats(self) -> Dict[str, int]:
        a = np.zeros(len(STATS_FIELDS), dtype=np.uint64)
        lib().vgo_get_stats(self.h, _ptr(a))
        return dict(zip(STATS_FIELDS, (int(x) for x in a)))

    def sites(self) -> np.ndarray:
        n = lib().vgo_get_sites(self.h, None, 0)
        out = np.zeros(n, dtype=SITE)
        lib().vgo_get_sites(self.h, _ptr(out), n)
        return out

    def add_counts(self, sites: np.ndarray):
        s = np.ascontiguousarray(sites, dtype=SITE)
        lib().vgo_add_counts(self.h, _ptr(s), s.size)

    # ---- stage F + G ----
    def calls(self) -> List[Tuple[str, int, int, float]]:
        """(chrom name from .chrlens, contig-relative POS, gtype, confidence) per called site, position order
        (qv.cc:1573-1626)."""
        out = []
        for s in self.sites():
            if s["ref"] == s["alt"]:
                continue
            gt, conf = call(int(s["ref_cnt"]), int(s["alt_cnt"]), int(s["ref_freq"]), int(s["alt_freq"]))
            if gt == 0:
                continue
            name, idx = chr_coord(int(s["pos"]), self.ix.chr_names, self.ix.chr_lens)
            out.append((name, idx, gt, conf))
        return out

    def write_vcf(self, vcf_in: str, vcf_out: str):
        write_vcf(self.calls(), vcf_in, vcf_out)


def call(ref_cnt: int, alt_cnt: int, ref_freq: int, alt_freq: int) -> Tuple[int, float]:
    conf = C.c_double()
    gt = lib().vgo_call(ref_cnt, alt_cnt, ref_freq, alt_freq, C.byref(conf))
    return gt, conf.value


def gq(conf: float) -> int:
    return lib().vgo_gq(conf)


def tables() -> Tuple[np.ndarray, np.ndarray]:
    g = np.zeros((64, 64, 3), dtype=np.float64)
    p = np.zeros(127, dtype=np.float64)
    lib().vgo_tables(_ptr(g), _ptr(p))
    return g, p


def ctx_digest(list_id: int, position: int, kmer_pos: int, kmer: int, modified: int) -> int:
    return lib().vgo_ctx_digest(list_id, position, kmer_pos, kmer, modified)


def chr_coord(index: int, names: List[str], lens: List[int]) -> Tuple[str, int]:
    """qv.cc:1590-1594 (names come from .chrlens, cut at 32 chars by qv.cc:488)."""
    j = 0
    while j < len(names) and index > lens[j]:
        index -= lens[j]
        j += 1
    return names[j], index


def write_vcf(calls, vcf_in: str, vcf_out: str) -> None:
    """Oracle-side restatement of the VCF rewrite, qv.cc:1628-1747."""
    table = {}
    for name, idx, gt, conf in calls:
        table["%s$%d" % (name, idx)] = (gt, conf)          # later duplicates overwrite, as unordered_map[] does
    has_gt = has_gq = False
    gt_index = gq_index = -1
    head_has_gt_col = True
    out = []
    with open(vcf_in, "r", newline="") as f:
        text = f.read()
    for line in text.split("\n"):                            # std::getline: '\r' stays in the line
        if line == "":
            continue
        if line[0] == "#" and line[1:2] == "#":
            out.append(line)
            if "ID=GT," in line:
                has_gt = True
            elif "ID=GQ," in line:
                has_gq = True
            continue
        if line[0] == "#":
            if not has_gt:
                out.append('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">')
                gt_index = 0
            if not has_gq:
                out.append('##FORMAT=<ID=GQ,Number=1,Type=Integer,Description="Genotype Quality">')
                gq_index = 1
            if len(line.split("\t")) < 10:
                head_has_gt_col = False
                line += "\tFORMAT\tDONOR"
            out.append(line)
            continue
        cols = line.split("\t")
        chrom = cols[0]
        if chrom[:1] != "c":
            chrom = "chr" + chrom
        key = chrom + "$" + cols[1]
        if key not in table:
            continue
        gt, conf = table[key]
        gts = GT_TEXT[gt]
        gqv = gq(conf)
        fmt = cols[8].split(":") if head_has_gt_col else []
        inf = cols[9].split(":") if head_has_gt_col else []
        if gt_index == -1 and has_gt:
            gt_index = fmt.index("GT")                      # assert(gt_index >= 0) in the reference
        if gt_index == -1 and has_gq:
            raise ValueError("header declares GQ but not GT: undefined in the reference (qv.cc:1699-1716)")
        if has_gt:
            inf[gt_index] = gts
        else:
            fmt.append("GT")
            inf.append(gts)
        if has_gq:
            inf[gq_index] = str(gqv)
        else:
            fmt.append("GQ")
            inf.append(str(gqv))
        if head_has_gt_col:
            cols[8] = ":".join(fmt)
            cols[9] = ":".join(inf)
        else:
            cols.append(":".join(fmt))
            cols.append(":".join(inf))
        out.append("\t".join(cols))
    with open(vcf_out, "w", newline="") as f:
        f.write("".join(l + "\n" for l in out))


# ---- the compiled reference (oracle/_ref), when present ----
def have_ref() -> bool:
    return os.path.exists(REF_BIN) and os.path.exists(REF_INSTR_BIN)


def run_ref_index(fasta: str, vcf: str, prefix: str) -> None:
    subprocess.check_call([REF_BIN, "index", fasta, vcf, prefix], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run_ref_geno(prefix: str, fastq: str, vcf: str, out_vcf: str, trace: Optional[str] = None,
                 dump: Optional[str] = None, instrumented: bool = True) -> float:
    """Runs the reference `geno`; returns its own 'Time:' figure (CPU seconds incl. load, qv.cc:1749-1751)."""
    env = dict(os.environ)
    if trace:
        env["VG_TRACE"] = trace
    if dump:
        env["VG_DUMP"] = dump
    exe = REF_INSTR_BIN if instrumented else REF_BIN
    p = subprocess.run([exe, "geno", prefix, fastq, vcf, out_vcf], env=env, stdout=subprocess.PIPE,
                       stderr=subprocess.DEVNULL, check=True, text=True)
    for line in p.stdout.splitlines():
        if line.startswith("Time:"):
            return float(line.split()[1])
    return math.nan


def parse_dump(path: str):
    """VG_DUMP file -> (sites SITE[], calls list)."""
    sites, calls = [], []
    for line in open(path):
        t = line.split()
        if t[0] == "P":
            sites.append((int(t[1]), int(t[2]), int(t[3]), int(t[4]), int(t[5]), int(t[6]), int(t[7]), 0, 0))
        elif t[0] == "C":
            calls.append((t[2], int(t[3]), int(t[4]), float(t[5])))
    return np.array(sites, dtype=SITE), calls
