"""Synthetic workloads built ON THE GPU (genome text, index, reads): what makes the GRCh38-shaped configurations of
BASELINE.json (configs[2..4]) feasible on a fresh box in seconds instead of the hour the
reference `index` needs.  Genome-sized byte arrays also live on the host (SNP selection, donor haplotypes: a few GB);
everything k-mer-sized (tens of GB) stays in HBM: vgb_build_index_device -> vgb_index_upload_device.

The index records produced this way are the ones `vargeno index` would write for the same FASTA + VCF
(tests/test_gpu_index_build.py: byte-identical on the golden sets, and equal to the numpy builder on a workload made here).
"""
from __future__ import annotations

import dataclasses
import time
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .. import geno
from . import index_builder as ib
from . import synth

U64 = np.uint64

# GRCh38 primary assembly chromosome lengths, chr1..chr22, chrX, chrY
GRCH38 = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259),
          ("chr6", 170805979), ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422),
          ("chr11", 135086622), ("chr12", 133275309), ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189),
          ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285), ("chr19", 58617616), ("chr20", 64444167),
          ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


@dataclasses.dataclass
class DeviceWorkload:
    name: str
    names: List[str]
    starts: np.ndarray
    lens: np.ndarray
    genome_len: int
    hap0_d: int
    hap1_d: int
    read_len: int
    seed: int
    n_snp_lines: int
    index_counts: dict
    setup_s: dict
    host_genome: Optional[np.ndarray] = None
    host_index: Optional[ib.Index] = None
    host_haps: Optional[Tuple[np.ndarray, np.ndarray]] = None
    host_snps: Optional[synth.SnpSet] = None


def repeat_ops(total: int, seed: int, frac: float) -> List[Tuple[int, int, int]]:
    """(src, dst, len) copies planting repeat families: ~frac of the sequence ends up being a copy (2..20 copies per family,
    500..2500 bases), so aux rows (2..10 copies) and POS_AMBIGUOUS (> 10) both occur."""
    ops, covered, fam = [], 0, 0
    while covered < frac * total:
        r = int(synth.rnd64(seed, 60, fam))
        ln = 500 + r % 2000
        copies = 2 + (r >> 20) % 19
        src = (r >> 28) % max(1, total - ln)
        for c in range(copies):
            d = int(synth.rnd64(seed, 61, fam, c)) % max(1, total - ln)
            if abs(d - int(src)) < ln:          # a device-to-device copy must not overlap its source
                continue
            ops.append((int(src), d, ln))
            covered += ln
        fam += 1
    return ops


def n_blocks(starts: np.ndarray, lens: np.ndarray, frac: float) -> List[Tuple[int, int]]:
    """(global start, length): one block at the start of every contig (telomere-like) and one in the middle, `frac` in total."""
    out = []
    for s, l in zip(starts, lens):
        a = int(l * frac * 0.6)
        b = int(l * frac * 0.4)
        if a:
            out.append((int(s), a))
        if b:
            out.append((int(s + l // 2), b))
    return out


def apply_layout_host(cat: np.ndarray, ops, blocks) -> None:
    """numpy twin of what build() does on the device (tests)."""
    for src, dst, ln in ops:
        cat[dst:dst + ln] = cat[src:src + ln].copy()
    for s, l in blocks:
        cat[s:s + l] = ord("N")


def _dict_side_ok(cat: np.ndarray, g: np.ndarray, chunk: int = 1 << 20) -> np.ndarray:
    """No N in [pos-32, pos+31] (src/dictgen.c:756-765), windows gathered in chunks to bound memory."""
    ok = np.zeros(g.size, bool)
    off = np.arange(-32, 32, dtype=np.int64)[None, :]
    for a in range(0, g.size, chunk):
        w = cat[g[a:a + chunk, None] + off]
        ok[a:a + chunk] = ~np.any(w == ord("N"), axis=1)
    return ok


def build(g: "geno.Genotyper", contigs: Sequence[Tuple[str, int]], n_snps: int, seed: int, name: str, n_frac: float = 0.05,
          repeat_frac: float = 0.02, read_len: int = 150, keep_host: bool = False, verbose: bool = False,
          blocks: Optional[List[Tuple[int, int]]] = None, motifs: Optional[List[Tuple[int, bytes]]] = None) -> DeviceWorkload:
    """blocks: explicit (global start, length) N runs instead of the n_frac layout.
    motifs: (global start, bytes) written over the random text after the repeat copies: a 16-base motif planted >= 100 times
    gives a reference HI32 block of >= 100 entries, i.e. the reference's "big" neighbour mode (src/qv.cc:242-264,962)."""
    t = {}
    t0 = time.time()
    names = [n for n, _ in contigs]
    lens = np.array([l for _, l in contigs], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
    total = int(lens.sum())
    gd = g.dalloc(total)
    geno.synth_genome_device(g, gd, starts, lens, seed)
    ops = repeat_ops(total, seed, repeat_frac)
    for src, dst, ln in ops:
        g._ck(g.L.vgb_memcpy_d2d(g.h, gd + dst, gd + src, ln))
    for pos, text in (motifs or []):
        g.h2d(gd + int(pos), np.frombuffer(text, dtype=np.uint8))
    if blocks is None:
        blocks = n_blocks(starts, lens, n_frac)
    for s, l in blocks:
        g._ck(g.L.vgb_memset_device(g.h, gd + s, ord("N"), l))
    t["genome_s"] = time.time() - t0
    t0 = time.time()
    cat = g.d2h(gd, total)
    gobj = synth.Genome(names, [cat[s:s + l] for s, l in zip(starts, lens)])
    snps = synth.make_snps(gobj, n_snps, seed=seed)
    gpos = starts[snps.contig] + snps.pos0
    dict_ok = _dict_side_ok(cat, gpos)
    f1 = np.round(1.0 - snps.caf_ref, 6)
    rf = ((f1.astype(np.float32) * np.float32(255.0)).astype(np.int64) & 0xFF).astype(np.uint8)
    af = ((snps.caf_ref.astype(np.float32) * np.float32(255.0)).astype(np.int64) & 0xFF).astype(np.uint8)
    code = synth._CODE
    h0, h1 = synth.donor_haplotypes(gobj, snps, seed=seed)
    t["snps_haps_s"] = time.time() - t0
    t0 = time.time()
    dix = geno.build_index_device(g, gd, names, starts, lens, gpos[dict_ok], code[snps.ref][dict_ok], code[snps.alt][dict_ok],
                                  rf[dict_ok], af[dict_ok], gpos)
    t["index_build_s"] = time.time() - t0
    counts = {"ref_kmers": int(dix.view.n_ref), "ref_aux_rows": int(dix.view.n_ref_aux), "snp_kmers": int(dix.view.n_snp),
              "snp_aux_rows": int(dix.view.n_snp_aux)}
    host_index = dix.to_host() if keep_host else None
    t0 = time.time()
    g.dfree(gd)
    geno.upload_device_index(g, dix)
    dix.free()
    t["index_upload_s"] = time.time() - t0
    h0d, h1d = g.dalloc(total), g.dalloc(total)
    g.h2d(h0d, h0)
    g.h2d(h1d, h1)
    if verbose:
        print("device workload %s: %s %s" % (name, counts, {k: round(v, 2) for k, v in t.items()}), flush=True)
    return DeviceWorkload(name, names, starts, lens, total, h0d, h1d, read_len, seed, int(dict_ok.sum()), counts, t,
                          cat if keep_host else None, host_index, (h0, h1) if keep_host else None, snps if keep_host else None)


def build_s1(g: "geno.Genotyper", scale: float = 1.0, seed: int = 7, keep_host: bool = False, verbose: bool = False) -> DeviceWorkload:
    """The S1 workload of tools/workloads.make_s1 (same genome, same SNP list, same index, byte for byte), built through the
    GPU in a few seconds instead of ~50 s of numpy."""
    L = int(50_800_000 * scale)
    n_snps = int(1_000_000 * scale)
    blocks = [(0, int(L * 0.19)), (int(L * 0.45), int(L * 0.02)), (int(L * 0.93), int(L * 0.006))]
    return build(g, [("chr22", L)], n_snps, seed, "S1 chr22-shaped x%g" % scale, repeat_frac=0.0, keep_host=keep_host, verbose=verbose,
                 blocks=blocks)


def synth_batch(g: "geno.Genotyper", wl: DeviceWorkload, out_d: int, n_reads: int, first_id: int, sub_rate: float, lowq_prob: float,
                lowq_chars: int = 4, id_width: int = 9) -> int:
    rb = 2 + id_width + 1 + wl.read_len + 3 + wl.read_len + 1
    g.synth_reads_device(wl.hap0_d, wl.hap1_d, wl.genome_len, wl.starts, wl.lens, n_reads, wl.read_len, wl.seed + 1000, first_id, id_width,
                         sub_rate, lowq_prob, lowq_chars, out_d, n_reads * rb)
    return n_reads * rb
