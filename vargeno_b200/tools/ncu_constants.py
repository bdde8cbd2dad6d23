"""profiles/ncu_constants.json from `ncu --page raw --csv` exports of the dominant kernel (one file per workload):

    python -m vargeno_b200.tools.ncu_constants s2=profiles/r02_k_geno8_s2_ncu_full.csv s3=... [--reads-per-launch 2000000]

bench.py reads the result (roofline.traffic, l2_requests_per_read, dram_fetches_per_read) and ignores it when the hash of the
kernel sources recorded here differs from the sources it runs with -- a profiler number must never outlive the kernel it
was taken from."""
from __future__ import annotations

import argparse
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    best = None
    for r in rows[2:]:
        m = {h: (v, u) for h, u, v in zip(hdr, units, r)}
        if best is None or float(m["gpu__time_duration.sum"][0]) > float(best["gpu__time_duration.sum"][0]):
            best = m
    return best


def num(m, key, scale_units=None):
    v, u = m[key]
    x = float(v)
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(u, 1.0)
    return x * mult


def main():
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("files", nargs="+", help="workload=csv")
    ap.add_argument("--reads-per-launch", type=int, default=2_000_000)
    args = ap.parse_args()
    out_path = os.path.join(ROOT, "profiles", "ncu_constants.json")
    out = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for spec in args.files:
        wl, path = spec.split("=", 1)
        m = metrics(path)
        n = args.reads_per_launch
        dram = num(m, "dram__bytes_read.sum") + num(m, "dram__bytes_write.sum")
        out[wl] = {
            "source": os.path.relpath(path, ROOT), "kernel": m["Kernel Name"][0], "kernel_hash": bench.kernel_hash(), "reads_per_launch": n,
            "launch_ms_under_ncu": num(m, "gpu__time_duration.sum") * 1e3,
            "dram_bytes_per_read": dram / n,
            "l2_requests_per_read": num(m, "lts__t_requests_srcunit_tex_op_read.sum") / n,
            # a miss of a .L2::64B load fetches two 32-byte sectors: one DRAM fetch per two missed sectors
            "dram_fetches_per_read": num(m, "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum") / 2 / n,
            "l2_hit_rate_pct": num(m, "lts__t_sector_hit_rate.pct"), "l1_hit_rate_pct": num(m, "l1tex__t_sector_hit_rate.pct"),
            "issue_active_pct": num(m, "smsp__issue_active.avg.pct_of_peak_sustained_active"), "registers": num(m, "launch__registers_per_thread"),
        }
    json.dump(out, open(out_path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
