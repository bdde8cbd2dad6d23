"""GRCh38-shaped run on one GPU (BASELINE.json configs[2], [3] and [4]):

    python -m vargeno_b200.tools.wgs_run [--scale 1.0] [--snps 12000000] [--reads 100000000] [--batch-reads 4000000]

  * 24 contigs with the GRCh38 chromosome lengths (x scale), 5 % N, 2 % planted repeats, 12 M SNPs (x scale);
    genome text, index and reads are all produced on the device (tools/device_workloads.py);
  * S2: 150 bp reads at 0.5 % substitutions, leading qualities low with p = 0.25; S3: 2 % substitutions, all leading
    qualities low; reads are generated batch by batch and never leave HBM;
  * S4: the dictionary-probe microbenchmark against the same index;
  * size-independent checks at this scale: chunking invariance and the shard-sum rule on one batch, and a truth check of
    the dictionary (k-mers cut from the genome at known positions must come back with exactly that position).
Prints one JSON object per section.
"""
from __future__ import annotations

import argparse
import json
import time

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--snps", type=int, default=12_000_000)
    ap.add_argument("--reads", type=int, default=100_000_000)
    ap.add_argument("--batch-reads", type=int, default=4_000_000)
    ap.add_argument("--probe-n", type=int, default=1 << 28)
    ap.add_argument("--skip-checks", action="store_true")
    args = ap.parse_args()
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import device_workloads as dw

    contigs = [(n, max(64, int(l * args.scale))) for n, l in dw.GRCH38]
    n_snps = max(100, int(args.snps * args.scale))
    B = args.batch_reads
    rb = 2 + 9 + 1 + 150 + 3 + 150 + 1
    g = Genotyper(device=0, max_chunk_bytes=B * rb + 4096)
    t0 = time.time()
    wl = dw.build(g, contigs, n_snps, seed=38, name="S2 GRCh38-shaped x%g" % args.scale, verbose=True, keep_host=not args.skip_checks)
    print(json.dumps({"section": "setup", "workload": wl.name, "genome_bp": wl.genome_len, "snp_lines_in_dict": wl.n_snp_lines,
                      "sites": g.n_sites, "index": wl.index_counts, "seconds": {k: round(v, 2) for k, v in wl.setup_s.items()},
                      "total_s": round(time.time() - t0, 1)}), flush=True)

    d = g.dalloc(2 * B * rb)                        # two batch buffers: generate one while the other is processed

    def run(tag, sub, lowq, n_reads):
        g.reset()
        nb = max(1, n_reads // B)
        # warm-up batch
        dw.synth_batch(g, wl, d, B, 0, sub, lowq)
        g.submit_device(d, B * rb, 0)
        g.sync()
        g.reset()
        t = time.perf_counter()
        for i in range(nb):
            buf = d + (i & 1) * B * rb
            dw.synth_batch(g, wl, buf, B, (i + 1) * B, sub, lowq)       # generation is part of the wall time reported below
            g.submit_device(buf, B * rb, (i + 1) * B)
        g.sync()
        wall = time.perf_counter() - t
        st = g.stats()
        gt, conf = g.call()
        look = st["exact_lookups"] + st["nbr_query_lookups"] + st["nbr_scan_reads"]
        kms = st["gpu_ms_geno"] + st["gpu_ms_parse"]
        print(json.dumps({"section": tag, "reads": st["reads"], "placed_fraction": st["placed"] / max(1, st["reads"]),
                          "reads_per_s_kernels": st["reads"] / (kms * 1e-3), "k_geno_ms_per_batch": st["gpu_ms_geno"] / nb,
                          "framing_ms_per_batch": st["gpu_ms_parse"] / nb, "lookups_per_read": look / max(1, st["reads"]),
                          "kmer_lookups_per_s": look / (st["gpu_ms_geno"] * 1e-3), "algorithmic_gbs": look * 32 / (st["gpu_ms_geno"] * 1e-3) / 1e9,
                          "wall_s_incl_read_generation": wall, "sites_called": int(np.count_nonzero(gt)), "sites": int(gt.size),
                          "ref_hom_het_alt": [int(np.count_nonzero(gt == k)) for k in (1, 3, 2)]}), flush=True)

    run("S2 reads 150bp 0.5% subst lowq 0.25", 0.005, 0.25, args.reads)
    run("S3 stress 2% subst all leading qualities low", 0.02, 1.0, max(B, args.reads // 4))

    rs = g.random_sector_bench(32 << 30, 1 << 30, 3)
    for mode, name in ((0, "uniform random 32-mers"), (1, "sampled reference-dictionary 32-mers"), (2, "half / half")):
        ms, found = g.probe_bench(args.probe_n, mode, seed=11, repeats=3)
        print(json.dumps({"section": "S4 probe microbench", "probe_set": name, "kmers_per_launch": args.probe_n,
                          "lookups_per_s": 2 * args.probe_n / (ms * 1e-3), "found_per_launch": found,
                          "algorithmic_gbs": 2 * args.probe_n * 32 / (ms * 1e-3) / 1e9, "random_sector_peak_gbs": rs,
                          "frac_of_random_sector_peak": 2 * args.probe_n * 32 / (ms * 1e-3) / 1e9 / rs}), flush=True)

    if not args.skip_checks:
        # (1) chunking invariance + shard-sum rule on one batch
        n = B
        dw.synth_batch(g, wl, d, n, 7 * B, 0.005, 0.25)
        def counters():
            p, m = g.counter_device_ptr()
            return g.d2h(p, m * 4).view(np.uint32).copy()
        g.reset(); g.submit_device(d, n * rb); g.sync(); whole = counters(); st_whole = g.stats()
        g.reset()
        cuts = [0, 3, 1001, n // 3, n // 3 + 1, n]
        for a, b in zip(cuts, cuts[1:]):
            g.submit_device(d + a * rb, (b - a) * rb, a)
        g.sync()
        same_chunks = bool(np.array_equal(counters(), whole)) and g.stats()["exact_lookups"] == st_whole["exact_lookups"]
        g.reset(); g.submit_device(d, (n // 2) * rb); g.sync(); a1 = counters()
        g.reset(); g.submit_device(d + (n // 2) * rb, (n - n // 2) * rb); g.sync(); a2 = counters()
        shard_sum = bool(np.array_equal(a1 + a2, whole))
        # (2) dictionary truth check: k-mers cut from the genome at known positions
        cat = wl.host_genome
        rng = np.random.default_rng(1)
        pos = rng.integers(0, wl.genome_len - 32, 200000)
        ci = np.searchsorted(wl.starts, pos, side="right") - 1
        pos = pos[pos + 32 <= wl.starts[ci] + wl.lens[ci]]
        win = cat[pos[:, None] + np.arange(32)[None, :]]
        okw = ~np.any(win == ord("N"), axis=1)
        pos, win = pos[okw], win[okw]
        from vargeno_b200.tools.synth import _CODE
        codes = _CODE[win].astype(np.uint64)
        km = np.zeros(pos.size, dtype=np.uint64)
        for b in range(32):
            km |= codes[:, b] << np.uint64(2 * b)
        hits = g.lookup(km)
        found = hits["ref_found"] == 1
        unamb = found & (hits["ref_flag"] == 0)
        exact_pos = bool(np.all(hits["ref_pos"][unamb] == (pos[unamb] + 1).astype(np.uint32)))
        # ambiguous ones: the true position must be one of the aux columns
        amb = np.flatnonzero(found & (hits["ref_flag"] == 1) & (hits["ref_pos"] != 0xFFFFFFFF))
        aux_ok = True
        if amb.size:
            aux = wl.host_index.ref_aux
            aux_ok = bool(np.all(np.any(aux[hits["ref_pos"][amb]] == (pos[amb] + 1)[:, None].astype(np.uint32), axis=1)))
        print(json.dumps({"section": "checks", "chunking_invariant": same_chunks, "shard_sum_rule": shard_sum,
                          "genome_kmers_probed": int(pos.size), "all_found": bool(found.all()), "unambiguous_positions_exact": exact_pos,
                          "ambiguous_positions_in_aux_row": aux_ok, "ambiguous_probed": int(amb.size),
                          "too_many_copies": int(np.count_nonzero(found & (hits["ref_pos"] == 0xFFFFFFFF)))}), flush=True)
    g.close()


if __name__ == "__main__":
    main()
