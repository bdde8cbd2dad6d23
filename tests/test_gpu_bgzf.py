"""SURVEY 8(f)-3 on the device: BGZF FASTQ input inflated by k_inflate_bgzf (one warp per gzip member, csrc/vgb_inflate.cuh --
the same decoder tests/test_inflate_host.py checks against zlib on the CPU) and framed with the overlap window
(k_fq_finish): every record must be processed exactly once, in whichever chunk it ENDS, whatever the member sizes and the chunk
size.  Compared read by read with the oracle, and through the command line with the golden VCF of the compiled reference."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from vargeno_b200 import build as vb
from vargeno_b200.tools import bgzf

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("layout", ["bgzip_64k", "tiny_and_uneven_members", "one_member_per_300_bytes"])
@pytest.mark.parametrize("own_out", [40_000, 1_000_000, 1 << 26])
def test_bgzf_chunks_process_every_read_once(cache, layout, own_out):
    from vargeno_b200.geno import Genotyper
    ix = cache.index("advA")                       # mixed read lengths 31..250, N and lower case in reads
    text = open(cache.dataset("advA").fastq, "rb").read()
    sizes = {"bgzip_64k": (), "tiny_and_uneven_members": (1000, 65280, 17, 30000, 1, 4096), "one_member_per_300_bytes": (300,)}[layout]
    if layout == "one_member_per_300_bytes":
        text = text[:1_200_000]
        text = text[:text.rfind(b"\n@") + 1]       # whole records
    buf = bgzf.compress(text, sizes=sizes, level=1 if layout != "bgzip_64k" else 6)
    o = orc.Oracle(ix)
    want = o.process_fastq(np.frombuffer(text, dtype=np.uint8))
    with Genotyper(device=0, trace=True, max_chunk_bytes=max(1 << 22, own_out + (1 << 18))) as g:
        g.upload_index(ix)
        n_chunks = g.submit_bgzf_file(buf, own_out)
        g.sync()
        got = g.read_results()
        r, a = g.pileup()
        st = g.stats()
    if own_out < len(text):
        assert n_chunks > 1
    assert got.size == want.size
    for f in ("flags", "freq", "n_ref", "n_snp", "passes", "ctx_hash"):
        bad = np.flatnonzero(got[f] != want[f])
        assert bad.size == 0, "%s differs for %d reads, first %s" % (f, bad.size, bad[:5])
    sites = o.sites()
    assert np.array_equal(r, sites["ref_cnt"]) and np.array_equal(a, sites["alt_cnt"])
    assert st["reads"] == want.size
    o.close()


def test_corrupt_member_and_truncated_tail_are_loud(cache):
    from vargeno_b200._lib import VgbError
    from vargeno_b200.geno import Genotyper
    ix = cache.index("s0")
    text = open(cache.dataset("s0").fastq, "rb").read()[:400_000]
    text = text[:text.rfind(b"\n@") + 1]
    with Genotyper(device=0, max_chunk_bytes=1 << 22) as g:
        g.upload_index(ix)
        buf = bytearray(bgzf.compress(text))
        buf[5000] ^= 0x55                          # flip bits inside the first member's DEFLATE payload
        g.submit_bgzf_file(bytes(buf), 1 << 26)
        with pytest.raises(VgbError):
            g.sync()
    with Genotyper(device=0, max_chunk_bytes=1 << 22) as g:
        g.upload_index(ix)
        g.submit_bgzf_file(bgzf.compress(text[:-155]), 1 << 26)     # the last record has lost its quality line: 3 lines left over
        with pytest.raises(VgbError):
            g.sync()


@pytest.mark.parametrize("chunk_mb", [1, 64])
def test_cli_bgzf_input_gives_the_reference_vcf(cache, tmp_path, chunk_mb):
    """`vargeno-b200 geno` on a .fastq.gz written as BGZF (and on two BGZF files back to back, cut inside a record)."""
    from vargeno_b200.tools import index_builder as ib
    vb.build()
    ds = cache.dataset("advA")
    prefix = str(tmp_path / "ix")
    ib.write_index(cache.index("advA"), prefix)
    text = open(ds.fastq, "rb").read()
    whole = tmp_path / "reads.fastq.gz"
    whole.write_bytes(bgzf.compress(text))
    out = str(tmp_path / "out.vcf")
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, str(whole), ds.vcf, out, "--chunk-mb", str(chunk_mb), "--verbose"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, "advA.out.vcf"), "rb").read()
    cut = len(text) // 2 + 23
    a, b = tmp_path / "a.fq.gz", tmp_path / "b.fq.gz"
    a.write_bytes(bgzf.compress(text[:cut]))
    b.write_bytes(bgzf.compress(text[cut:], sizes=(5000, 65280)))
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, "%s,%s" % (a, b), ds.vcf, out, "--chunk-mb", str(chunk_mb)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, "advA.out.vcf"), "rb").read()
