"""Host half of stage F + stage G of the C++ program (csrc/host/geno_host.cpp): concatenated position -> contig coordinate
(src/qv.cc:1590-1594), GQ = (int)(-10 log(conf)) (:1681) and the VCF rewrite (:1628-1747) must reproduce the VCFs the compiled
reference wrote (tests/golden/*.out.vcf) when fed the per-site calls of the golden runs.  No GPU: `vargeno-b200 vcf-rewrite`."""
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc
from vargeno_b200 import build as vb

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_vcf_rewrite_matches_reference(cache, name, tmp_path):
    vb.build()
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    ix = cache.index(name)
    prefix = str(tmp_path / "ix")
    with open(prefix + ".chrlens", "w") as f:
        for n, l in zip(ix.chr_names, ix.chr_lens):
            f.write("%s %d\n" % (n, l))
    # one line per static-pileup site, position order, as vgb_fetch_sites + vgb_call deliver them
    calls = str(tmp_path / "calls.tsv")
    n_called = 0
    with open(calls, "w") as f:
        for s in gold["sites"]:
            gt, conf = (0, 0.0)
            if s["ref"] != s["alt"]:
                gt, conf = orc.call(int(s["ref_cnt"]), int(s["alt_cnt"]), int(s["ref_freq"]), int(s["alt_freq"]))
            n_called += gt != 0
            f.write("%d %d %s\n" % (int(s["pos"]), gt, float(conf).hex()))
    assert n_called == gold["call_gt"].size
    out = str(tmp_path / "out.vcf")
    p = subprocess.run([vb.HOST_BIN, "vcf-rewrite", prefix, calls, cache.dataset(name).vcf, out], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".out.vcf"), "rb").read()


def test_vcf_rewrite_usage(tmp_path):
    vb.build()
    p = subprocess.run([vb.HOST_BIN, "vcf-rewrite", "too", "few"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert p.returncode != 0
