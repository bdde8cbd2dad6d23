"""Dictionary-probe microbenchmark (BASELINE.json configs[4]) against the S1 index: device-generated 32-mers probed in both
dictionaries (query_ref_dict + query_snp_dict per k-mer), reported as lookups/s and as a fraction of the measured
random-sector rate of the same GPU.

    python -m vargeno_b200.tools.probe_bench [--scale 1.0] [--n 268435456] [--repeats 5]
"""
from __future__ import annotations

import argparse
import json


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--n", type=int, default=1 << 28)
    ap.add_argument("--repeats", type=int, default=5)
    args = ap.parse_args()
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import workloads
    wl = workloads.make_s1(scale=args.scale)
    with Genotyper(device=0) as g:
        g.upload_index(wl.index)
        rs = g.random_sector_bench(32 << 30, 1 << 30, 3)
        for mode, name in ((0, "uniform random 32-mers (misses)"), (1, "sampled reference-dictionary 32-mers (hits)"), (2, "half / half")):
            ms, found = g.probe_bench(args.n, mode, seed=11, repeats=args.repeats)
            lookups = 2 * args.n                      # one reference + one SNP dictionary query per k-mer
            print(json.dumps({"probe_set": name, "kmers_per_launch": args.n, "lookups_per_s": lookups / (ms * 1e-3), "ms_per_launch": ms,
                              "found_per_launch": found, "algorithmic_gbs": lookups * 32 / (ms * 1e-3) / 1e9,
                              "random_sector_peak_gbs": rs, "frac_of_random_sector_peak": lookups * 32 / (ms * 1e-3) / 1e9 / rs}), flush=True)


if __name__ == "__main__":
    main()
