import sys, os, time, subprocess
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from vargeno_b200.geno import Genotyper
from vargeno_b200.tools import workloads
from bench import rec_bytes, REC_ID_WIDTH, ClockSampler
wl = workloads.make_s1(scale=0.1)
B, L, nb = 2_000_000, wl.read_len, 6
bb = B * rec_bytes(L)
g = Genotyper(device=0, max_chunk_bytes=bb + 4096)
g.upload_index(wl.index)
h0, h1 = g.dalloc(wl.haps[0].size), g.dalloc(wl.haps[1].size)
g.h2d(h0, wl.haps[0]); g.h2d(h1, wl.haps[1])
d = g.dalloc(nb * bb)
g.synth_reads_device(h0, h1, wl.haps[0].size, wl.genome.starts, wl.genome.lengths, nb * B, L, 5, 0, REC_ID_WIDTH, 0.005, 0.25, 4, d, nb * bb)
pinned = torch.empty(nb * bb, dtype=torch.uint8, pin_memory=True)
host = pinned.numpy()
host[:] = g.d2h(d, nb * bb)
def leg(tag):
    g.reset()
    ts = []
    for rep in range(4):
        t = time.perf_counter()
        for i in range(nb):
            g.submit_chunk(host[i * bb:(i + 1) * bb])
        g.sync()
        ts.append((time.perf_counter() - t) / nb * 1e3)
    print(tag, " ".join("%.1f" % x for x in ts), "ms/step", flush=True)
leg("no sampler     ")
s = ClockSampler(0); s.start(); time.sleep(0.5)
leg("with nvidia-smi")
print(s.finish())
leg("after sampler  ")
# device-resident leg interleaved, then e2e again (the order bench.py uses)
for i in range(nb): g.submit_device(d + i * bb, bb)
g.sync()
leg("after resident ")
gt, cf = g.call()
leg("after call     ")
