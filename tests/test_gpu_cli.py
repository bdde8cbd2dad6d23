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


def test_cli_two_gpus(cache, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    got, err = _run(cache, "advA", tmp_path, ["--chunk-mb", "1", "--gpus", "2"])
    assert got == open(os.path.join(GOLD, "advA.out.vcf"), "rb").read()


def test_cli_rejects_bad_usage(tmp_path):
    vb.build()
    p = subprocess.run([vb.HOST_BIN, "geno", "only", "three", "args"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0
