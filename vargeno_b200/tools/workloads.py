"""The synthetic workloads BASELINE.json names, as pure functions of a seed."""
from __future__ import annotations

import dataclasses
from typing import Tuple

import numpy as np

from . import index_builder as ib
from . import synth


@dataclasses.dataclass
class Workload:
    name: str
    genome: synth.Genome
    snps: synth.SnpSet
    haps: Tuple[np.ndarray, np.ndarray]
    index: ib.Index
    read_len: int
    sub_rate: float
    lowq_prob: float
    lowq_chars: int
    seed: int


def caf_pair(snps: synth.SnpSet):
    """The two CAF numbers exactly as synth.write_vcf prints them."""
    return np.round(1.0 - snps.caf_ref, 6), snps.caf_ref


def make_s1(scale: float = 1.0, seed: int = 7, sub_rate: float = 0.005, lowq_prob: float = 0.25) -> Workload:
    """S1, chr22-shaped: one contig of 50.8 Mbp with ~11 Mbp of N in 3 blocks, 1 M bi-allelic SNPs, 150 bp reads at
    0.5 % substitutions, the first four quality characters below '8' with probability 0.25 each.
    scale < 1 shrinks genome and SNP count proportionally (tests)."""
    L = int(50_800_000 * scale)
    n_snps = int(1_000_000 * scale)
    n_blocks = [(0, 0, int(L * 0.19)), (0, int(L * 0.45), int(L * 0.02)), (0, int(L * 0.93), int(L * 0.006))]
    g = synth.make_genome([("chr22", L)], seed=seed, n_runs=n_blocks)
    snps = synth.make_snps(g, n_snps, seed=seed)
    haps = synth.donor_haplotypes(g, snps, seed=seed)
    f1, f2 = caf_pair(snps)
    ix = ib.build_index_from_arrays(g.names, g.seqs, snps.contig, snps.pos0, snps.ref, snps.alt, f1, f2)
    return Workload("S1 chr22-shaped x%g" % scale, g, snps, haps, ix, 150, sub_rate, lowq_prob, 4, seed)
