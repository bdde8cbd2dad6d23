"""GRCh38-SHAPED parity (BASELINE configs[2] / [3] = SURVEY.md 8(d) S2 / S3, scaled down so that the CPU oracle finishes in
seconds): 24 contigs with the GRCh38 length proportions, N blocks, planted repeat families (2..20 copies -> aux rows and
POS_AMBIGUOUS entries, src/dictgen.c:63-154) and a planted 16-base motif (reference HI32 block >= 100 entries -> the "big"
neighbour mode, src/qv.cc:242-264,962-1109).  Index built on the device (vgb_build_index_device), reads generated on the
device plus targeted reads over the motif sites and the repeat copies; then the CUDA path against the oracle: per-read flags,
vote, context digest, pileup counters, lookup statistics, genotype calls -- and the two-context shard-sum tail
(vgb_counter_device_ptr + sum + clamp + call) that a multi-GPU run takes.

Reference semantics covered: src/qv.cc:850-937 (exact contexts through aux rows), :962-1109 (big mode), :1110-1365."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

REC_ID_WIDTH = 9
SCALE = 0.02
MOTIF = b"GATTACAGGCTTCAGA"
KEYS = ("reads", "skipped_n", "passes", "placed", "exact_lookups", "nbr_query_lookups", "nbr_scan_reads", "bf_probes", "lowq_kmers",
        "events", "pileup_incr", "big_kmers")


def rec_bytes(L):
    return 2 + REC_ID_WIDTH + 1 + L + 3 + L + 1


class Shaped:
    pass


@pytest.fixture(scope="module")
def shaped():
    """The scaled GRCh38-shaped workload: host index + genome + haplotypes, and the motif / repeat bookkeeping."""
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import device_workloads as dw
    from vargeno_b200.tools import synth
    contigs = [(n, max(4096, int(l * SCALE))) for n, l in dw.GRCH38]
    total = sum(l for _, l in contigs)
    # 160 motif sites, away from contig ends (positions are global; the N layout of build() occupies the first 3 % and a
    # stretch after the middle of every contig: keep to the second quarter)
    starts = np.concatenate([[0], np.cumsum([l for _, l in contigs])[:-1]])
    sites = []
    for k in range(160):
        c = k % len(contigs)
        r = int(synth.rnd64(77, 5, k))
        sites.append(int(starts[c] + contigs[c][1] // 4 + r % (contigs[c][1] // 5)))
    sh = Shaped()
    with Genotyper(device=0) as g:
        wl = dw.build(g, contigs, int(12_000_000 * SCALE), seed=38, name="S2-shaped x%g" % SCALE, keep_host=True,
                      motifs=[(p, MOTIF) for p in sites])
        L = wl.read_len
        n2, n3 = 60_000, 24_000
        d = g.dalloc((n2 + n3) * rec_bytes(L))
        dw.synth_batch(g, wl, d, n2, 0, 0.005, 0.25)                                   # S2 style
        dw.synth_batch(g, wl, d + n2 * rec_bytes(L), n3, n2, 0.02, 1.0)                # S3 style
        sh.text2 = g.d2h(d, n2 * rec_bytes(L)).copy()
        sh.text3 = g.d2h(d + n2 * rec_bytes(L), n3 * rec_bytes(L)).copy()
        g.dfree(d)
        g.dfree(wl.hap0_d)
        g.dfree(wl.hap1_d)
    sh.index, sh.cat, sh.haps, sh.L = wl.host_index, wl.host_genome, wl.host_haps, L
    sh.names, sh.starts, sh.lens = wl.names, wl.starts, wl.lens
    gobj = synth.Genome(wl.names, [sh.cat[s:s + l] for s, l in zip(wl.starts, wl.lens)])
    # targeted reads: the motif on bases 16..31 of k-mer j of the read (both strands), every leading quality low
    ts, rv = [], []
    for k, p in enumerate(sites):
        for j in range(4):
            s = p - 16 - 32 * j
            ts.append(s)
            rv.append((k + j) % 3 == 0)
    ts = np.array(ts, dtype=np.int64)
    first = n2 + n3
    t_motif = synth.simulate_reads(gobj, sh.haps, ts.size, L, seed=91, sub_rate=0.02, lowq_prob=1.0, first_id=first, forced_starts=ts,
                                   forced_rev=np.array(rv))
    first += ts.size
    # reads over the planted repeat copies (sources and destinations): aux rows and POS_AMBIGUOUS entries get queried
    ops = dw.repeat_ops(total, 38, 0.02)
    rs = np.array([dst + (i * 13) % max(1, ln - L) for i, (src, dst, ln) in enumerate(ops) if ln >= L][:6000], dtype=np.int64)
    t_rep = synth.simulate_reads(gobj, sh.haps, rs.size, L, seed=93, sub_rate=0.01, lowq_prob=0.6, first_id=first, forced_starts=rs)
    sh.text_t = np.concatenate([t_motif, t_rep])
    return sh


def _run_gpu(index, text, trace=True, chunk=1 << 22):
    from vargeno_b200.geno import Genotyper
    with Genotyper(device=0, trace=trace, max_chunk_bytes=chunk) as g:
        g.upload_index(index)
        g.submit(text)
        g.sync()
        res = g.read_results() if trace else None
        r, a = g.pileup()
        gt, conf = g.call()
        st = g.stats()
    return res, r, a, gt, conf, st


def _check(index, text):
    o = orc.Oracle(index)
    want = o.process_fastq(text)
    res, r, a, gt, conf, st = _run_gpu(index, text)
    assert res.size == want.size
    for f in ("flags", "freq", "n_ref", "n_snp", "passes", "ctx_hash"):
        bad = np.flatnonzero(res[f] != want[f])
        assert bad.size == 0, "%s differs from the oracle for %d reads, first %s" % (f, bad.size, bad[:5])
    placed = (want["flags"] & orc.F_PROCESS) != 0
    assert np.array_equal(res["target"][placed], want["target"][placed])
    sites = o.sites()
    assert np.array_equal(r, sites["ref_cnt"]) and np.array_equal(a, sites["alt_cnt"])
    ost = o.stats()
    for k in KEYS:
        assert st[k] == ost[k], k
    exp = [orc.call(int(s["ref_cnt"]), int(s["alt_cnt"]), int(s["ref_freq"]), int(s["alt_freq"])) if s["ref"] != s["alt"] else (0, 0.0)
           for s in sites]
    assert np.array_equal(gt, np.array([e[0] for e in exp], np.uint8))
    assert np.array_equal(conf, np.array([e[1] for e in exp]))          # bit-identical, stricter than the 1e-9 the spec allows
    o.close()
    return ost, want


def test_index_has_the_shapes_that_matter(shaped):
    ix = shaped.index
    assert ix.ref_aux.shape[0] > 100                                    # 2..10-copy repeats -> aux rows
    assert np.count_nonzero(ix.ref["pos"] == 0xFFFFFFFF) > 100          # > 10 copies -> POS_AMBIGUOUS
    hi = (ix.ref["kmer"] >> np.uint64(32)).astype(np.uint32)            # sorted by k-mer, so equal HI32 values are adjacent
    edges = np.flatnonzero(np.diff(hi)) + 1
    assert np.diff(np.concatenate([[0], edges, [hi.size]])).max() >= 100  # a HI32 block of >= 100 entries: big mode exists


def test_s2_style_reads_match_the_oracle(shaped):
    ost, want = _check(shaped.index, shaped.text2)
    assert ost["reads"] == 60_000 and ost["placed"] > 30_000


def test_s3_style_reads_match_the_oracle(shaped):
    ost, want = _check(shaped.index, shaped.text3)
    assert ost["reads"] == 24_000 and ost["lowq_kmers"] >= 4 * 24_000   # every k-mer of every pass is neighbour-searched


@pytest.mark.parametrize("shortscan", ["0", "1"])
def test_both_scan_instantiations_match_the_oracle(shaped, shortscan, monkeypatch):
    """The strided-scan instantiation is chosen per index by the mean HI24 block length (VGB_SHORTSCAN overrides): a scaled-down
    index always gets the short one, the GRCh38-sized one of the benchmark the other -- both must give the oracle's answers."""
    monkeypatch.setenv("VGB_SHORTSCAN", shortscan)
    _check(shaped.index, shaped.text3)


def test_targeted_reads_fire_big_mode_aux_rows_and_ambiguity(shaped):
    ost, want = _check(shaped.index, shaped.text_t)
    assert ost["big_kmers"] > 100                                       # src/qv.cc:962: ref block >= 100
    assert np.count_nonzero(want["flags"] & orc.F_AMBIGUOUS) > 0        # two positions tie (repeat copies)
    assert np.count_nonzero(want["n_ref"] > 4) > 100                    # aux rows expanded into several contexts per k-mer


def test_two_contexts_shard_sum_clamp_and_call(shaped):
    """What N GPUs do, on the one GPU the test box has: two contexts with the same index take one read shard
    each; the raw counters of the second are fetched through vgb_counter_device_ptr and added to the first's (the job of
    vgb_allreduce_pileup), then clamp + caller run on the sum.  Must equal the oracle over all reads."""
    from vargeno_b200.geno import Genotyper
    text = np.concatenate([shaped.text2, shaped.text3, shaped.text_t])
    rb = rec_bytes(shaped.L)
    n = text.size // rb
    cut = (n * 2 // 5) * rb
    o = orc.Oracle(shaped.index)
    o.process_fastq(text, want_results=False)
    sites = o.sites()
    with Genotyper(device=0, max_chunk_bytes=1 << 23) as g0, Genotyper(device=0, max_chunk_bytes=1 << 23) as g1:
        g0.upload_index(shaped.index)
        g1.upload_index(shaped.index)
        g0.submit(text[:cut])
        g1.submit(text[cut:], first_read_id=cut // rb)
        g0.sync()
        g1.sync()
        p0, n0 = g0.counter_device_ptr()
        p1, n1 = g1.counter_device_ptr()
        assert n0 == n1 == 2 * g0.n_sites
        c0 = g0.d2h(p0, n0 * 4).view(np.uint32)
        c1 = g1.d2h(p1, n1 * 4).view(np.uint32)
        assert c0.any() and c1.any()
        g0.h2d(p0, (c0 + c1).astype(np.uint32))                         # the sum a ncclAllReduce would leave in place
        r, a = g0.pileup()
        gt, conf = g0.call()
        s0, s1 = g0.stats(), g1.stats()
    assert np.array_equal(r, sites["ref_cnt"]) and np.array_equal(a, sites["alt_cnt"])
    # saturation is applied after the sum (F10)
    assert np.array_equal(r, np.minimum(c0[0::2] + c1[0::2], 63)) and np.array_equal(a, np.minimum(c0[1::2] + c1[1::2], 63))
    exp = [orc.call(int(s["ref_cnt"]), int(s["alt_cnt"]), int(s["ref_freq"]), int(s["alt_freq"])) if s["ref"] != s["alt"] else (0, 0.0)
           for s in sites]
    assert np.array_equal(gt, np.array([e[0] for e in exp], np.uint8)) and np.array_equal(conf, np.array([e[1] for e in exp]))
    ost = o.stats()
    for k in KEYS:
        assert s0[k] + s1[k] == ost[k], k
    o.close()
