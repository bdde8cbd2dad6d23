"""The GPU index builder (vgb_build_index_device) must produce the records the compiled reference `vargeno index` wrote:
the records copied back from HBM are written with tools.index_builder.write_index and compared with the sha256 of the
reference-built files (tests/golden/*.json).  Also: uploading the device-resident records gives the same genotyping
results as uploading the host image, and the device genome generator is the twin of tools/synth.make_genome."""
import hashlib
import json
import os

import numpy as np
import pytest

from vargeno_b200.tools import index_builder as ib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


def _device_index(g, ds):
    from vargeno_b200 import geno
    raw_names, raw_seqs = ib.read_fasta_raw(ds.fasta)
    names, seqs = ib.normalise_fasta(raw_names, raw_seqs)
    starts = np.concatenate([[0], np.cumsum([s.size for s in seqs])[:-1]]).astype(np.int64)
    lines = ib.parse_vcf_for_dict(ds.vcf, names, seqs)
    bc, bp = ib.bf_line_positions(ds.vcf, raw_names, raw_seqs)
    cat = np.concatenate(seqs)
    gd = g.dalloc(cat.size)
    g.h2d(gd, cat)
    dix = geno.build_index_device(g, gd, names, starts, [s.size for s in seqs], starts[lines.contig] + lines.index0, lines.ref_code,
                                  lines.alt_code, lines.ref_freq, lines.alt_freq, starts[bc] + bp)
    g.dfree(gd)
    return dix


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_gpu_built_index_is_byte_identical(cache, name, tmp_path):
    from vargeno_b200.geno import Genotyper
    man = json.load(open(os.path.join(GOLD, name + ".json")))
    ds = cache.dataset(name)
    with Genotyper(device=0) as g:
        dix = _device_index(g, ds)
        host = dix.to_host()
        dix.free()
    prefix = str(tmp_path / "gpu")
    ib.write_index(host, prefix)
    for ext in ("ref.dict", "snp.dict", "ref.bf", "snp.bf", "chrlens"):
        assert os.path.getsize(prefix + "." + ext) == man["index_bytes"][ext], ext
        assert _sha(prefix + "." + ext) == man["index"][ext], ext


def test_device_resident_upload_equals_host_upload(cache):
    from vargeno_b200 import geno
    from vargeno_b200.geno import Genotyper
    ds = cache.dataset("advB")
    fq = np.fromfile(ds.fastq, dtype=np.uint8)
    with Genotyper(device=0, trace=True) as g:
        g.upload_index(cache.index("advB"))
        g.submit(fq)
        g.sync()
        want_reads, want_cnt, want_sites = g.read_results(), g.pileup(), g.sites()
    with Genotyper(device=0, trace=True) as g:
        dix = _device_index(g, ds)
        geno.upload_device_index(g, dix)
        dix.free()
        g.submit(fq)
        g.sync()
        got_reads, got_cnt, got_sites = g.read_results(), g.pileup(), g.sites()
    assert np.array_equal(want_reads, got_reads)
    assert np.array_equal(want_cnt[0], got_cnt[0]) and np.array_equal(want_cnt[1], got_cnt[1])
    for k in want_sites:
        assert np.array_equal(want_sites[k], got_sites[k]), k


def test_device_workload_matches_numpy_builder():
    """tools/device_workloads.build (genome layout, SNP filters and index all via the GPU) gives the index the numpy builder
    -- itself pinned to the reference -- gives for the same genome and SNP list."""
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import device_workloads as dw
    from vargeno_b200.tools import synth, workloads
    contigs = [("chr1", 300000), ("chr2", 180000), ("chrM", 16571)]
    with Genotyper(device=0) as g:
        wl = dw.build(g, contigs, 3000, seed=5, name="t", keep_host=True)
        assert g.n_sites > 2000
    g0 = synth.make_genome(contigs, seed=5)
    cat = g0.concat().copy()
    dw.apply_layout_host(cat, dw.repeat_ops(cat.size, 5, 0.02), dw.n_blocks(wl.starts, wl.lens, 0.05))
    assert np.array_equal(cat, wl.host_genome)
    seqs = [cat[s:s + l] for s, l in zip(wl.starts, wl.lens)]
    snps = synth.make_snps(synth.Genome(wl.names, seqs), 3000, seed=5)
    f1, f2 = workloads.caf_pair(snps)
    want = ib.build_index_from_arrays(wl.names, seqs, snps.contig, snps.pos0, snps.ref, snps.alt, f1, f2)
    got = wl.host_index
    assert want.ref_aux.shape[0] > 0, "repeat families should create aux rows"
    for f in ("ref", "ref_aux", "snp", "snp_aux", "snp_bf"):
        assert np.array_equal(getattr(want, f), getattr(got, f)), f
    assert np.array_equal(np.flatnonzero(want.ref_bf), np.flatnonzero(got.ref_bf[:want.ref_bf.size]))


def test_device_s1_equals_numpy_s1():
    """bench.py builds S1 through the device; it must be the S1 of tools/workloads.make_s1 (genome, haplotypes, index)."""
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import device_workloads as dw
    from vargeno_b200.tools import workloads
    want = workloads.make_s1(scale=0.004)
    with Genotyper(device=0) as g:
        wl = dw.build_s1(g, scale=0.004, keep_host=True)
        h0 = g.d2h(wl.hap0_d, wl.genome_len)
        h1 = g.d2h(wl.hap1_d, wl.genome_len)
    assert np.array_equal(wl.host_genome, want.genome.concat())
    assert np.array_equal(h0, want.haps[0]) and np.array_equal(h1, want.haps[1])
    for f in ("ref", "ref_aux", "snp", "snp_aux", "snp_bf"):
        assert np.array_equal(getattr(want.index, f), getattr(wl.host_index, f)), f
    assert np.array_equal(np.flatnonzero(want.index.ref_bf), np.flatnonzero(wl.host_index.ref_bf))


def test_device_genome_generator_matches_numpy():
    from vargeno_b200 import geno
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import synth
    g0 = synth.make_genome([("chrA", 70001), ("chrB", 12345), ("chrC", 64)], seed=99)
    with Genotyper(device=0) as g:
        total = sum(g0.lengths)
        d = g.dalloc(total)
        geno.synth_genome_device(g, d, g0.starts, g0.lengths, 99)
        got = g.d2h(d, total)
        g.dfree(d)
    assert np.array_equal(got, g0.concat())
