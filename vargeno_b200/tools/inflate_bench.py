"""Device-side BGZF inflate (k_inflate_bgzf) measured alone: S1-like FASTQ text made on the device, written as BGZF by zlib at
two levels, pushed through vgb_submit_bgzf from pinned memory in chunks; prints the rate at which text comes out of the inflate +
framing kernels (vgb_stats.gpu_ms_parse: CUDA events around both) and the wall-clock rate of the whole submit loop.

    python -m vargeno_b200.tools.inflate_bench [--reads 4000000] [--chunk-mb 256] [--levels 1,6]
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np


def _members(args):
    from vargeno_b200.tools import bgzf
    data, level, block = args
    return b"".join(bgzf.member(data[a:a + block], level) for a in range(0, len(data), block))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--chunk-mb", type=int, default=256, help="inflated bytes per vgb_submit_bgzf call")
    ap.add_argument("--levels", default="1,6")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    import torch

    import bench
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import bgzf
    from vargeno_b200.tools import device_workloads as dw

    rb = bench.rec_bytes()
    own = args.chunk_mb << 20
    with Genotyper(device=0, max_chunk_bytes=own + (1 << 20)) as g:
        wl = dw.build_s1(g, scale=0.25)
        d = g.dalloc(args.reads * rb)
        dw.synth_batch(g, wl, d, args.reads, 0, 0.005, 0.25, bench.LOWQ_CHARS, bench.REC_ID_WIDTH)
        text = g.d2h(d, args.reads * rb).tobytes()
        g.dfree(d)
        for level in (int(x) for x in args.levels.split(",")):
            t0 = time.time()
            span = 65280 * 256
            with mp.Pool(max(1, (os.cpu_count() or 2) - 1)) as pool:
                buf = b"".join(pool.imap(_members, [(text[a:a + span], level, 65280) for a in range(0, len(text), span)])) + bgzf.EOF_MEMBER
            t_comp = time.time() - t0
            mem, plan = bgzf.plan_chunks(bgzf.scan(buf), own)
            chunks = []
            for k, (o0, i0, i1, ov) in enumerate(plan):
                comp, tab = bgzf.chunk_arrays(buf, mem, o0, i1)
                pin = torch.empty(comp.size, dtype=torch.uint8, pin_memory=True).numpy()
                pin[:] = comp
                chunks.append((pin, tab, ov, k + 1 == len(plan)))
            best = None
            for rep in range(args.repeats + 1):
                g.reset()
                s0 = g.stats()
                t0 = time.perf_counter()
                for pin, tab, ov, last in chunks:
                    g.submit_bgzf(pin, tab, ov, last)
                g.sync()
                dt = time.perf_counter() - t0
                s1 = g.stats()
                ms_parse = s1["gpu_ms_parse"] - s0["gpu_ms_parse"]
                assert s1["reads"] - s0["reads"] == args.reads, (s1["reads"], args.reads)
                if rep and (best is None or ms_parse < best[0]):
                    best = (ms_parse, dt, s1["gpu_ms_geno"] - s0["gpu_ms_geno"])
            print(json.dumps({"zlib_level": level, "text_bytes": len(text), "bgzf_bytes": len(buf), "ratio": len(text) / len(buf), "members": len(mem),
                              "chunks": len(chunks), "compress_s": round(t_comp, 1), "inflate_plus_framing_ms": best[0], "geno_ms": best[2],
                              "inflate_text_gb_per_s": len(text) / (best[0] * 1e-3) / 1e9, "wall_s": best[1],
                              "wall_text_gb_per_s": len(text) / best[1] / 1e9, "wall_reads_per_s": args.reads / best[1]}), flush=True)


if __name__ == "__main__":
    main()
