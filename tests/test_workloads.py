import numpy as np

from vargeno_b200.tools import index_builder as ib
from vargeno_b200.tools import synth, workloads


def test_array_index_path_equals_file_path(tmp_path):
    """bench.py builds its index straight from arrays; it must be the index `vargeno index` would build from the files."""
    w = workloads.make_s1(scale=0.004, seed=7)
    fa, vcf = str(tmp_path / "r.fa"), str(tmp_path / "s.vcf")
    synth.write_fasta(w.genome, fa)
    synth.write_vcf(w.genome, w.snps, vcf)
    ref = ib.build_index(fa, vcf)
    for f in ("ref", "ref_aux", "snp", "snp_aux", "snp_bf"):
        assert np.array_equal(getattr(ref, f), getattr(w.index, f)), f
    assert np.array_equal(np.flatnonzero(ref.ref_bf), np.flatnonzero(w.index.ref_bf))
    assert ref.chr_names == w.index.chr_names and ref.chr_lens == w.index.chr_lens
    assert w.index.snp.size > 0 and (w.index.snp.size < 32 * w.snps.pos0.size)   # some SNPs fall next to the N blocks
