"""Pins the CPU oracle (oracle/vg_oracle.c) to outputs of the compiled reference (tests/golden/*.npz, *.out.vcf)
and to the known-answer vector of the reference's own test/expected_output."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def oracles(cache):
    made = {}

    def get(name):
        if name not in made:
            o = orc.Oracle(cache.index(name))
            fq = np.fromfile(cache.dataset(name).fastq, dtype=np.uint8)
            made[name] = (o, o.process_fastq(fq))
        return made[name]
    yield get
    for o, _ in made.values():
        o.close()


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_vote_trace_matches_reference(oracles, name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    _, res = oracles(name)
    g = gold["reads"]
    assert res.size == g.size
    for f in ("flags", "target", "freq", "n_ref", "n_snp", "ctx_hash"):
        bad = np.flatnonzero(res[f] != g[f])
        assert bad.size == 0, "%s differs for %d reads, first %s" % (f, bad.size, bad[:5])


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_pileup_and_calls_match_reference(oracles, name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    o, _ = oracles(name)
    assert np.array_equal(o.sites(), gold["sites"])
    calls = o.calls()
    assert [c[0] for c in calls] == list(gold["call_chr"])
    assert np.array_equal(np.array([c[1] for c in calls]), gold["call_pos"])
    assert np.array_equal(np.array([c[2] for c in calls], np.uint8), gold["call_gt"])
    # bit-exact: same glibc libm builds the tables, same IEEE operation order
    assert np.array_equal(np.array([c[3] for c in calls]), gold["call_conf"])


@pytest.mark.parametrize("name", ["s0", "advA", "advB"])
def test_vcf_matches_reference(oracles, cache, name, tmp_path):
    o, _ = oracles(name)
    out = str(tmp_path / "o.vcf")
    o.write_vcf(cache.dataset(name).vcf, out)
    assert open(out, "rb").read() == open(os.path.join(GOLD, name + ".out.vcf"), "rb").read()


def test_big_mode_and_retries_exercised(oracles):
    o, res = oracles("advB")
    st = o.stats()
    assert st["big_kmers"] > 100 and st["nbr_scan_reads"] > 0 and st["nbr_query_lookups"] > 0
    assert np.count_nonzero(res["flags"] & orc.F_REVCOMPL) > 0
    o, res = oracles("advA")
    assert np.count_nonzero(res["flags"] & orc.F_SKIPPED) > 0


def test_known_answer_expected_output():
    """reference test/expected_output:59-64: every record has GQ 846, i.e. 63 supporting reads of one allele.
    (63, 0) -> 0/0 with confidence 1.7723448369165199e-37."""
    gt, conf = orc.call(63, 0, int(np.float32(0.9992) * np.float32(255)), int(np.float32(0.0007987) * np.float32(255)))
    assert gt == 1 and orc.gq(conf) == 846
    # the survey's figure came from a Python restatement (different libm): agree to 1e-12, and exactly with glibc's table
    assert abs(conf - 1.7723448369165199e-37) / conf < 1e-12
    assert conf == orc.tables()[1][63]
    gt, conf = orc.call(0, 63, 253, 1)
    assert gt == 2 and orc.gq(conf) == 846
    assert orc.call(0, 0, 127, 127)[0] == 0 and orc.call(63, 63, 127, 127)[0] == 0


def test_counter_reduction_is_shard_invariant(cache):
    """per-shard counters combined with min(63, a+b) equal the single-run counters."""
    ix = cache.index("s0")
    fq = np.fromfile(cache.dataset("s0").fastq, dtype=np.uint8)
    nl = np.flatnonzero(fq == 10)
    cut = int(nl[4 * 7001 - 1]) + 1
    whole, a, b = orc.Oracle(ix), orc.Oracle(ix), orc.Oracle(ix)
    whole.process_fastq(fq, want_results=False)
    a.process_fastq(fq[:cut], want_results=False)
    b.process_fastq(fq[cut:], want_results=False)
    a.add_counts(b.sites())
    assert np.array_equal(a.sites(), whole.sites())
    for o in (whole, a, b):
        o.close()


@pytest.mark.ref
@pytest.mark.parametrize("name", ["s0"])
def test_rerun_compiled_reference(cache, name, tmp_path):
    """Opt-in (VG_RUN_REF=1): run oracle/_ref again and diff the full trace text, not just the digests."""
    assert orc.have_ref()
    from vargeno_b200.tools import index_builder as ib
    ds = cache.dataset(name)
    prefix = str(tmp_path / "ref")
    ib.write_index(cache.index(name), prefix)
    orc.run_ref_geno(prefix, ds.fastq, ds.vcf, str(tmp_path / "out.vcf"), trace=str(tmp_path / "t_ref"), dump=str(tmp_path / "d"))
    o = orc.Oracle(cache.index(name))
    o.process_fastq(np.fromfile(ds.fastq, dtype=np.uint8), trace_path=str(tmp_path / "t_mine"))
    assert open(str(tmp_path / "t_ref"), "rb").read() == open(str(tmp_path / "t_mine"), "rb").read()


@pytest.mark.ref
def test_rerun_reference_index_big100(cache, tmp_path):
    """Opt-in (VG_RUN_REF=1, ~2 min, ~6 GB): the compiled reference's `vargeno index` on the 105 Mbp repeat-family set writes the
    files whose sha256 tests/golden/big100.json holds (what `vargeno-b200 index` is compared with on the GPU box)."""
    import hashlib
    import json
    import subprocess
    assert orc.have_ref()
    man = json.load(open(os.path.join(GOLD, "big100.json")))
    ds = cache.dataset("big100")
    prefix = str(tmp_path / "ref")
    subprocess.check_call([orc.REF_BIN, "index", ds.fasta, ds.vcf, prefix], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    for ext, want in man["index"].items():
        h = hashlib.sha256()
        with open(prefix + "." + ext, "rb") as f:
            for blk in iter(lambda: f.read(1 << 24), b""):
                h.update(blk)
        assert h.hexdigest() == want, ext
