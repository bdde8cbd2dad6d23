"""The C++ host program `vargeno-b200 geno` must write the VCF the compiled reference wrote (tests/golden/*.out.vcf),
byte for byte, from index files in the reference's own on-disk format."""
import os
import subprocess

import pytest

from vargeno_b200 import build as vb
from vargeno_b200.tools import index_builder as ib

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _run(cache, name, tmp_path, extra):
    vb.build()
    ds = cache.dataset(name)
    prefix = str(tmp_path / "ix")
    ib.write_index(cache.index(name), prefix)
    out = str(tmp_path / "out.vcf")
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, ds.fastq, ds.vcf, out, "--verbose"] + extra, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    return open(out, "rb").read(), p.stderr


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_cli_vcf_is_byte_identical(cache, name, tmp_path):
    got, _ = _run(cache, name, tmp_path, ["--chunk-mb", "1"])
    assert got == open(os.path.join(GOLD, name + ".out.vcf"), "rb").read()


def test_cli_default_chunking(cache, tmp_path):
    got, err = _run(cache, "s0", tmp_path, [])
    assert got == open(os.path.join(GOLD, "s0.out.vcf"), "rb").read()
    assert '"reads": 20000' in err


def _gpu_count():
    """GPUs on the box, asked from the driver's own tool (no framework import inside the test process)."""
    try:
        out = subprocess.run(["nvidia-smi", "-L"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=60).stdout
    except Exception:
        return 0
    return sum(1 for l in out.splitlines() if l.startswith("GPU "))


def test_cli_two_gpus(cache, tmp_path):
    if _gpu_count() < 2:
        pytest.skip("needs 2 GPUs")
    got, err = _run(cache, "advA", tmp_path, ["--chunk-mb", "1", "--gpus", "2"])
    assert got == open(os.path.join(GOLD, "advA.out.vcf"), "rb").read()


def test_cli_rejects_bad_usage(tmp_path):
    vb.build()
    p = subprocess.run([vb.HOST_BIN, "geno", "only", "three", "args"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0


def _sha(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_cli_index_files_are_byte_identical(cache, name, tmp_path):
    """`vargeno-b200 index` (text parsing on the host, sort / collapse / Bloom filters on the GPU) must write the files the
    compiled reference's `vargeno index` wrote -- the five `geno` reads and the .ref.bf.lite.bf nothing reads
    (src/generate_bf.cc:102-105,145-163): sizes and sha256 from tests/golden/<name>.json."""
    import json
    vb.build()
    man = json.load(open(os.path.join(GOLD, name + ".json")))
    ds = cache.dataset(name)
    prefix = str(tmp_path / "cli")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.vcf, prefix, "--verbose"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    for ext in ("ref.dict", "snp.dict", "ref.bf", "snp.bf", "chrlens", "ref.bf.lite.bf"):
        assert os.path.getsize(prefix + "." + ext) == man["index_bytes"][ext], ext
        assert _sha(prefix + "." + ext) == man["index"][ext], ext
    if name == "s0":            # --no-lite leaves the sixth file out and changes nothing else
        p2 = str(tmp_path / "nolite")
        p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.vcf, p2, "--no-lite"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert p.returncode == 0, p.stderr
        assert not os.path.exists(p2 + ".ref.bf.lite.bf") and _sha(p2 + ".ref.dict") == man["index"]["ref.dict"]


def test_cli_index_at_100_mbp_with_repeat_families(cache, tmp_path):
    """SURVEY 8(f)-1 at a size where the reference's `qsort` order of EQUAL k-mers shows (aux-row column order): 105 Mbp, 400
    repeat families, 686 589 aux rows.  tests/golden/big100.json holds the sha256 of what the compiled reference's `vargeno
    index` wrote for exactly this FASTA + VCF (inputs are pinned by their own sha256; the reference run is reproducible with
    VG_RUN_REF=1 pytest tests/test_oracle_golden.py -k big100)."""
    import json
    vb.build()
    man = json.load(open(os.path.join(GOLD, "big100.json")))
    ds = cache.dataset("big100")
    assert _sha(ds.fasta) == man["inputs"]["ref.fa"] and _sha(ds.vcf) == man["inputs"]["snp.vcf"]
    prefix = str(tmp_path / "big")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.vcf, prefix, "--verbose"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    for ext in ("chrlens", "snp.bf", "snp.dict", "ref.dict", "ref.bf", "ref.bf.lite.bf"):
        assert os.path.getsize(prefix + "." + ext) == man["index_bytes"][ext], ext
        assert _sha(prefix + "." + ext) == man["index"][ext], ext


def test_cli_index_from_ucsc_table_then_geno(cache, tmp_path):
    """`vargeno-b200 index ref.fa snps.txt prefix` (UCSC snp-table input) writes the files the compiled reference's
    `ucscd` + `ucscbf` wrote, and `geno` on that index gives the reference's VCF (tests/golden/ucscA.*)."""
    import json
    vb.build()
    man = json.load(open(os.path.join(GOLD, "ucscA.json")))
    ds = cache.dataset("ucscA")
    prefix = str(tmp_path / "u")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.txt, prefix], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    for ext in ("ref.dict", "snp.dict", "ref.bf", "snp.bf", "chrlens", "ref.bf.lite.bf"):
        assert os.path.getsize(prefix + "." + ext) == man["index_bytes"][ext], ext
        assert _sha(prefix + "." + ext) == man["index"][ext], ext
    out = str(tmp_path / "out.vcf")
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, ds.fastq, ds.vcf, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, "ucscA.out.vcf"), "rb").read()


def test_cli_index_then_geno(cache, tmp_path):
    """The two commands back to back, as a user of the reference would run them."""
    vb.build()
    ds = cache.dataset("s0")
    prefix = str(tmp_path / "ix")
    p = subprocess.run([vb.HOST_BIN, "index", ds.fasta, ds.vcf, prefix, "--no-lite"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    out = str(tmp_path / "out.vcf")
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, ds.fastq, ds.vcf, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, "s0.out.vcf"), "rb").read()


def test_cli_gzip_and_split_fastq(cache, tmp_path):
    """gzip input and a comma-separated list of files read back to back (mates of a paired run are concatenated
    for the reference, experiment/experiment.md:22-27).  The split is NOT at a record boundary and one half is gzip: the VCF
    must still be the reference's, byte for byte."""
    import gzip
    vb.build()
    ds = cache.dataset("advA")
    prefix = str(tmp_path / "ix")
    ib.write_index(cache.index("advA"), prefix)
    text = open(ds.fastq, "rb").read()
    cut = len(text) // 2 + 17
    a, b = str(tmp_path / "a.fq.gz"), str(tmp_path / "b.fq")
    with gzip.open(a, "wb", compresslevel=1) as f:
        f.write(text[:cut])
    open(b, "wb").write(text[cut:])
    out = str(tmp_path / "out.vcf")
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, a + "," + b, ds.vcf, out, "--chunk-mb", "1"], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, "advA.out.vcf"), "rb").read()
    whole = str(tmp_path / "whole.fq.gz")
    with gzip.open(whole, "wb", compresslevel=1) as f:
        f.write(text)
    p = subprocess.run([vb.HOST_BIN, "geno", prefix, whole, ds.vcf, out], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, "advA.out.vcf"), "rb").read()
