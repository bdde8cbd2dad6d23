"""Size-independent properties at bench scale (millions of reads, device-generated), where the oracle is too slow to
check everything: chunking invariance, shard-sum invariance (the multi-GPU reduction rule), determinism, plus oracle
parity on a subsample -- for the S1 shape (BASELINE configs[1]), the high-error stress shape (configs[3] at S1 scale)
and the probe microbenchmark (configs[4])."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

REC_ID_WIDTH = 9


def rec_bytes(L):
    return 2 + REC_ID_WIDTH + 1 + L + 3 + L + 1


@pytest.fixture(scope="module")
def s1_quarter():
    from vargeno_b200.tools import workloads
    return workloads.make_s1(scale=0.25)


def _device_reads(g, wl, n, seed, sub_rate, lowq_prob):
    L = wl.read_len
    h0, h1 = g.dalloc(wl.haps[0].size), g.dalloc(wl.haps[1].size)
    g.h2d(h0, wl.haps[0])
    g.h2d(h1, wl.haps[1])
    d = g.dalloc(n * rec_bytes(L))
    g.synth_reads_device(h0, h1, wl.haps[0].size, wl.genome.starts, wl.genome.lengths, n, L, seed, 0, REC_ID_WIDTH, sub_rate,
                         lowq_prob, wl.lowq_chars, d, n * rec_bytes(L))
    g.dfree(h0)
    g.dfree(h1)
    return d


def _raw_counters(g):
    p, n = g.counter_device_ptr()
    return g.d2h(p, n * 4).view(np.uint32).copy()


KEYS = ("reads", "skipped_n", "passes", "placed", "exact_lookups", "nbr_query_lookups", "nbr_scan_reads", "bf_probes", "lowq_kmers",
        "events", "pileup_incr", "big_kmers")


@pytest.mark.parametrize("shape", ["s1", "stress"])
def test_fullsize_properties(s1_quarter, shape):
    from vargeno_b200.geno import Genotyper
    wl = s1_quarter
    sub, lowq = (0.005, 0.25) if shape == "s1" else (0.02, 1.0)     # stress: 2 % substitutions, every leading quality low
    n = 3_000_000
    rb = rec_bytes(wl.read_len)
    with Genotyper(device=0, max_chunk_bytes=n * rb + 4096) as g:
        g.upload_index(wl.index)
        d = _device_reads(g, wl, n, 4242, sub, lowq)
        # one chunk
        g.submit_device(d, n * rb)
        g.sync()
        whole, st_whole = _raw_counters(g), g.stats()
        gt_whole, conf_whole = g.call()
        assert st_whole["reads"] == n and st_whole["placed"] > 0.5 * n * (0.6 if shape == "stress" else 1.0)
        # determinism
        g.reset()
        g.submit_device(d, n * rb)
        g.sync()
        assert np.array_equal(_raw_counters(g), whole)
        # uneven chunks: same counters, same statistics
        g.reset()
        cuts = [0, 1, 17, 4096, 1_000_003, 1_000_004, 2_500_000, n]
        for a, b in zip(cuts, cuts[1:]):
            g.submit_device(d + a * rb, (b - a) * rb, a)
        g.sync()
        assert np.array_equal(_raw_counters(g), whole)
        st = g.stats()
        for k in KEYS:
            assert st[k] == st_whole[k], k
        # shard-sum rule: counters of two read shards add up exactly; calls after the sum are the same
        g.reset()
        g.submit_device(d, 1_234_567 * rb)
        g.sync()
        first = _raw_counters(g)
        g.reset()
        g.submit_device(d + 1_234_567 * rb, (n - 1_234_567) * rb, 1_234_567)
        g.sync()
        second = _raw_counters(g)
        assert np.array_equal(first + second, whole)
        # the host copy of the first reads agrees with the oracle read by read
        m = 20000
        text = g.d2h(d, m * rb)
        g.dfree(d)
    o = orc.Oracle(wl.index)
    want = o.process_fastq(text)
    with Genotyper(device=0, trace=True, max_chunk_bytes=m * rb + 4096) as g:
        g.upload_index(wl.index)
        g.submit_chunk(text)
        g.sync()
        got = g.read_results()
        r, a = g.pileup()
        st = g.stats()
    for f in ("flags", "freq", "n_ref", "n_snp", "passes", "ctx_hash"):
        assert np.array_equal(got[f], want[f]), f
    sites = o.sites()
    assert np.array_equal(r, sites["ref_cnt"]) and np.array_equal(a, sites["alt_cnt"])
    ost = o.stats()
    for k in KEYS:
        assert st[k] == ost[k], k
    if shape == "stress":
        assert ost["lowq_kmers"] >= 4 * m          # every k-mer of every pass is neighbour-searched
    o.close()
    assert gt_whole.size == conf_whole.size and np.count_nonzero(gt_whole) > 0.5 * gt_whole.size


def test_probe_microbench_counts(s1_quarter):
    """configs[4] at S1 scale: device-generated probe batches; sampled dictionary k-mers are all found, random ones are not."""
    from vargeno_b200.geno import Genotyper
    wl = s1_quarter
    with Genotyper(device=0) as g:
        g.upload_index(wl.index)
        n = 1 << 22
        ms_hit, found_hit = g.probe_bench(n, 1, seed=3, repeats=2)
        ms_miss, found_miss = g.probe_bench(n, 0, seed=3, repeats=2)
        ms_mix, found_mix = g.probe_bench(n, 2, seed=3, repeats=2)
    assert found_hit >= n                         # every sampled reference k-mer is found (a few are in the SNP dictionary too)
    assert found_hit < n + n // 10
    assert found_miss <= 4                        # 2^22 random 64-bit keys against ~10^7 entries: essentially never
    assert n // 2 <= found_mix <= n // 2 + n // 10
    assert ms_hit > 0 and ms_miss > 0 and ms_mix > 0
