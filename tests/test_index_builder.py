"""The numpy index builder must reproduce the files the compiled reference `vargeno index` wrote
(sha256 recorded in tests/golden/<name>.json by tests/golden/make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from vargeno_b200.tools import index_builder as ib

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_inputs_and_index_match_reference(cache, name):
    man = json.load(open(os.path.join(GOLD, name + ".json")))
    ds = cache.dataset(name)
    assert _sha(ds.fasta) == man["inputs"]["fasta"]
    assert _sha(ds.vcf) == man["inputs"]["vcf"]
    assert _sha(ds.fastq) == man["inputs"]["fastq"]
    ix = cache.index(name)
    prefix = os.path.join(ds.dir, "mine")
    ib.write_index(ix, prefix)
    try:
        for ext in ("ref.dict", "snp.dict", "ref.bf", "snp.bf", "chrlens"):
            assert os.path.getsize(prefix + "." + ext) == man["index_bytes"][ext], ext
            assert _sha(prefix + "." + ext) == man["index"][ext], ext
        # and the reader gives back what was written
        back = ib.load_index(prefix)
        assert np.array_equal(back.ref, ix.ref) and np.array_equal(back.ref_aux, ix.ref_aux)
        assert np.array_equal(back.snp, ix.snp) and np.array_equal(back.snp_aux, ix.snp_aux)
        assert np.array_equal(back.snp_bf, ix.snp_bf) and back.chr_names == ix.chr_names and back.chr_lens == ix.chr_lens
        nz = np.flatnonzero(ix.ref_bf)
        assert np.array_equal(back.ref_bf[nz], ix.ref_bf[nz]) and np.count_nonzero(back.ref_bf) == nz.size
    finally:
        for ext in ("ref.dict", "snp.dict", "ref.bf", "snp.bf", "chrlens"):
            if os.path.exists(prefix + "." + ext):
                os.remove(prefix + "." + ext)


def test_adversarial_features_present(cache):
    """The adversarial sets must really contain what they claim to exercise."""
    a = cache.index("advA")
    b = cache.index("advB")
    assert a.ref_aux.shape[0] > 0 and np.any(a.ref["pos"][a.ref["flag"] == 1] == 0xFFFFFFFF)
    assert b.snp_aux.size > 0, "SNP aux rows (ambiguous SNP k-mers)"
    assert np.any((b.snp["flag"] == 1) & (b.snp["pos"] == 0xFFFFFFFF)), "POS_AMBIGUOUS in the SNP dictionary"
    hi = (b.ref["kmer"] >> np.uint64(32)).astype(np.uint32)
    _, counts = np.unique(hi, return_counts=True)
    assert counts.max() >= 100, "a ref HI32 block >= BLOCK_SIZE_THRESHOLD"
    assert hi.max() == 0xFFFFFFFF, "last jumpgate block populated"


def test_lite_filter_matches_reference(cache):
    """The unused .ref.bf.lite.bf (2.3 GB) is written by `index` too; check it once, on the smoke set."""
    man = json.load(open(os.path.join(GOLD, "s0.json")))
    ds = cache.dataset("s0")
    names, seqs = ib.read_fasta_raw(ds.fasta)
    pck = [ib.contig_kmers(s) for s in seqs]
    _, lite = ib.build_ref_bf(pck, want_lite=True)
    h = hashlib.sha256()
    h.update(np.array([ib.REF_LITE_BF_BITS], dtype="<u8").tobytes())
    h.update(memoryview(np.ascontiguousarray(lite)))
    assert h.hexdigest() == man["index"]["ref.bf.lite.bf"]
