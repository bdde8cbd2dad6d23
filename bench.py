#!/usr/bin/env python
"""bench.py -- reads/s (and k-mer lookups/s) of the `vargeno geno` hot path on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W]              our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]        the reference's own CPU geno on the host cores

Workload (BASELINE.json configs[1], SURVEY.md 8(d) "S1"): synthetic chr22-shaped reference (50.8 Mbp, ~11 Mbp N),
1 M-SNP list, 150 bp reads at 0.5 % substitutions, first four quality characters below '8' with probability 0.25.
One step = one batch of `--batch-reads` FASTQ records (default 2 M = 630 MB of text) through the whole per-read path
(framing, 2-bit packing, exact + Hamming-1 lookups, Bloom gates, vote, pileup atomics).  Every step gets a distinct
batch that is larger than the 126 MB L2, so no L2 flush is needed between steps.

Timed regions are bracketed by a barrier and a device synchronise on both sides (wall clock, max over ranks); the
kernel time for the roofline comes from CUDA events recorded on the launching stream inside libvgb200.so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REC_ID_WIDTH = 9
# DRAM bytes one k_geno8 launch really moves (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of
# this workload at 2 M reads per launch: profiles/r01_k_geno8_ncu_full.csv); a number taken under the profiler, so it is a
# constant here, not something measured in the timed run
NCU_TRAFFIC = {"bytes": 3.120863e9 + 55.968512e6, "reads_per_launch": 2_000_000, "source": "profiles/r01_k_geno8_ncu_full.csv"}
S1_SUB_RATE, S1_LOWQ_PROB, S1_LOWQ_CHARS = 0.005, 0.25, 4


def rec_bytes(read_len):
    return 2 + REC_ID_WIDTH + 1 + read_len + 3 + read_len + 1


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
        except Exception:
            pass
    return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device = device
        self.rows = []
        self.stop_flag = threading.Event()
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag.is_set():
                    break
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def finish(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_near_gpu(device):
    """Run this rank on the CPUs NVML names as closest to its GPU, so that the pinned FASTQ buffers (first touched by this
    process) sit in host memory of the same NUMA node as the GPU's PCIe root.  Returns what was done, for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
        cpus = [c for c in cpus if c < ncpu]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "cpus %d-%d" % (min(cpus), max(cpus)) if len(cpus) > 1 else "cpu %d" % cpus[0]
    except Exception as e:     # best effort: the number is still valid without it, only possibly slower
        return "unbound (%s)" % type(e).__name__
    return "unbound"


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import device_workloads as dw

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_near_gpu(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    K, W, B = args.steps, args.warmup, args.batch_reads
    t_setup = time.time()
    L = 150
    batch_bytes = B * rec_bytes(L)
    nb = K + W

    uid = None
    if world > 1:
        box = [Genotyper.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]
    g = Genotyper(device=local, max_chunk_bytes=batch_bytes + 4096, world_size=world, rank=rank, nccl_unique_id=uid)
    # S1 index: genome text, SNP list, reference-format index records and the GPU re-layout, all produced through the device
    # (vgb_build_index_device is byte-identical to `vargeno index`: tests/test_gpu_index_build.py); replicated on every rank
    want_cpu = rank == 0 and world == 1 and not args.skip_cpu and args.workload == "s1"
    if args.workload == "s1":
        wl = dw.build_s1(g, scale=args.scale, keep_host=want_cpu)
        sub_rate, lowq_prob, shape = S1_SUB_RATE, S1_LOWQ_PROB, "1M-SNP list, 150bp reads @0.5% subst, lowq 0.25 on first 4 quality chars"
    else:
        # BASELINE.json configs[2] / [3]: GRCh38-shaped reference, 12 M SNPs; ~89 GiB of HBM per GPU.
        # The CPU baseline leg is skipped here (the oracle would need the 37 GB index image on the host).
        contigs = [(n, max(64, int(l * args.scale))) for n, l in dw.GRCH38]
        wl = dw.build(g, contigs, max(100, int(12_000_000 * args.scale)), seed=38, name="S2 GRCh38-shaped x%g" % args.scale)
        sub_rate, lowq_prob = (0.005, 0.25) if args.workload == "s2" else (0.02, 1.0)
        shape = "12M-SNP list, 150bp reads @%g%% subst, lowq %g on first 4 quality chars" % (sub_rate * 100, lowq_prob)
    if world > 1:
        g.allreduce()                       # NCCL connection set-up happens on the first collective: keep it out of the timed legs

    # synthetic reads, generated on the device (byte-identical twin of tools/synth.simulate_reads)
    d_reads = g.dalloc(nb * batch_bytes)
    first = rank * nb * B
    dw.synth_batch(g, wl, d_reads, nb * B, first, sub_rate, lowq_prob, S1_LOWQ_CHARS, REC_ID_WIDTH)
    # host copy in pinned memory for the end-to-end leg
    pinned = torch.empty(nb * batch_bytes, dtype=torch.uint8, pin_memory=True)
    host = pinned.numpy()
    host[:] = g.d2h(d_reads, nb * batch_bytes)
    # pinned destination of the job's result (GT + confidence per SNP site), allocated once like a real caller would
    out_gt = torch.empty(g.n_sites, dtype=torch.uint8, pin_memory=True).numpy()
    out_conf = torch.empty(g.n_sites, dtype=torch.float64, pin_memory=True).numpy()
    setup_s = time.time() - t_setup

    def barrier():
        g.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def lookups(st):
        return st["exact_lookups"] + st["nbr_query_lookups"] + st["nbr_scan_reads"]

    # ---- leg 1: inputs resident in HBM ----
    g.reset()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.5)                         # nvidia-smi needs a moment to come up; samples cover both timed legs
    for i in range(W):
        g.submit_device(d_reads + i * batch_bytes, batch_bytes, first + i * B)
    barrier()
    st0 = g.stats()
    t0 = time.perf_counter()
    for i in range(W, nb):
        g.submit_device(d_reads + i * batch_bytes, batch_bytes, first + i * B)
    g.sync()
    if world > 1:
        g.allreduce()                       # the one exchange step of the job (NCCL sum of the per-SNP counters)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    st1 = g.stats()
    n_sites = g.n_sites
    reads_step = B * world
    value = K * reads_step / dt
    d_lookups = lookups(st1) - lookups(st0)
    d_ms_geno = st1["gpu_ms_geno"] - st0["gpu_ms_geno"]
    d_ms_parse = st1["gpu_ms_parse"] - st0["gpu_ms_parse"]
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    peak, peak_src = measured_peaks()
    alg_bytes = d_lookups * 32.0
    achieved = alg_bytes / (d_ms_geno * 1e-3) / 1e9 if d_ms_geno > 0 else 0.0

    # ---- leg 2: end to end through the C ABI with host buffers (H2D inside the timed region, calls read back) ----
    g.reset()
    for i in range(W):
        g.submit_chunk(host[i * batch_bytes:(i + 1) * batch_bytes], first + i * B)
    g.sync()
    g.call(out=(out_gt, out_conf))          # warm-up of the result path too (its device staging is allocated at first use)
    barrier()
    t0 = time.perf_counter()
    for i in range(W, nb):
        g.submit_chunk(host[i * batch_bytes:(i + 1) * batch_bytes], first + i * B)
    g.sync()
    t_reads = time.perf_counter() - t0
    if world > 1:
        g.allreduce()
    t_reduce = time.perf_counter() - t0 - t_reads
    gt, conf = g.call(out=(out_gt, out_conf))   # device -> host read of the job's result (GT + confidence per SNP site)
    t_call = time.perf_counter() - t0 - t_reads - t_reduce
    barrier()
    dt_e2e = max_over_ranks(time.perf_counter() - t0)
    e2e_value = K * reads_step / dt_e2e
    # keep the GPU busy a little longer so that the 100 ms clock sampler sees the part under this load
    t_end = time.perf_counter() + args.clock_hold_s
    while time.perf_counter() < t_end:
        for i in range(nb):
            g.submit_device(d_reads + i * batch_bytes, batch_bytes, first + i * B)
        g.sync()
    clocks = sampler.finish()

    out = None
    if rank == 0:
        rs = None
        if world == 1 and not args.skip_roofline_probe:
            try:
                rs = g.random_sector_bench(32 << 30, 1 << 30, 3)
            except Exception:
                rs = None
        cpu = None
        if want_cpu:
            cpu = cpu_port_baseline(wl.host_index, host[:min(B, args.cpu_sample) * rec_bytes(L)])
        out = {
            "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic",
            "config": {"workload": wl.name + ", " + shape,
                       "reads_per_step_per_gpu": B, "read_len": L, "fastq_bytes_per_step_per_gpu": batch_bytes,
                       "index": "replicated per GPU", "reads": "sharded across GPUs",
                       "l2": "every step streams a distinct batch larger than L2 (no flush needed)",
                       "ref_kmers": wl.index_counts["ref_kmers"], "snp_kmers": wl.index_counts["snp_kmers"], "snp_sites": int(n_sites),
                       "index_built_by": "vgb_build_index_device (byte-identical to `vargeno index`)"},
            "kmer_lookups_per_s": d_lookups * world / dt if world == 1 else None,
            "lookups_per_read": d_lookups / max(1, st1["reads"] - st0["reads"]),
            "placed_fraction": (st1["placed"] - st0["placed"]) / max(1, st1["reads"] - st0["reads"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_TRAFFIC["bytes"] * B / NCU_TRAFFIC["reads_per_launch"] if (args.scale == 1.0 and args.workload == "s1") else None,
                         "traffic_source": NCU_TRAFFIC["source"],
                         "kernel": "k_geno8 (+ list-mode k_geno for deferred reads)", "algorithmic_bytes_per_launch": alg_bytes / K, "launch_ms": d_ms_geno / K,
                         "peak_source": peak_src, "note": "32 B (one DRAM sector) per dictionary lookup, SURVEY.md 8(d)",
                         "random_sector_peak_gbs": rs, "frac_of_random_sector_peak": (achieved / rs) if rs else None},
            "kernel_ms_per_step": {"k_geno": d_ms_geno / K, "fastq_framing": d_ms_parse / K},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": batch_bytes * world,
                    "d2h_bytes_per_step": int(n_sites * 9 * world / K), "ms_per_step": dt_e2e / K * 1e3,
                    "rank0_ms": {"submit_and_sync": t_reads * 1e3, "allreduce": t_reduce * 1e3, "call_d2h": t_call * 1e3}},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "host_binding": numa,
            "setup_s": setup_s,
        }
        emit(out)
    g.dfree(d_reads)
    g.close()
    if world > 1:
        dist.destroy_process_group()
    return out


def cpu_port_baseline(index, text):
    """Bounded single-thread run of the CPU oracle (kind "port") over a prefix of the same reads."""
    from oracle import oracle as orc
    o = orc.Oracle(index)
    t0 = time.perf_counter()
    o.process_fastq(np.ascontiguousarray(text), want_results=False)
    dt = time.perf_counter() - t0
    st = o.stats()
    o.close()
    return {"value": st["reads"] / dt, "unit": "reads/s", "cores": 1, "kind": "port",
            "sample": "first %d reads of step 0, oracle/liboracle.so single thread, index resident, %.1f s" % (st["reads"], dt),
            "kmer_lookups_per_s": (st["exact_lookups"] + st["nbr_query_lookups"] + st["nbr_scan_reads"]) / dt}


# ------------------------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference binary, one process per host core that fits in RAM, fed through FIFOs
# ------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import tempfile

    from oracle import oracle as orc
    from vargeno_b200.tools import index_builder as ib
    from vargeno_b200.tools import synth, workloads

    K, W = args.steps, args.warmup
    L = 150
    Bp = args.ref_batch_reads
    try:
        avail_gb = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) / 1e6
    except Exception:
        avail_gb = 32
    use_ref = orc.have_ref() and avail_gb > 40
    nproc = max(1, min(os.cpu_count() or 1, int((avail_gb - 16) // 19))) if use_ref else 1
    if args.ref_procs:
        nproc = args.ref_procs
    nb = K + W

    # index + reads (distinct per process and step).  With a GPU on the box they come from the device-side builder (seconds,
    # byte-identical index); without one, from the numpy tooling.  Neither is part of what is timed.
    total = nproc * nb * Bp
    index, text, name = None, None, None
    try:
        import torch
        if torch.cuda.is_available():
            from vargeno_b200.geno import Genotyper
            from vargeno_b200.tools import device_workloads as dw
            with Genotyper(device=0) as g:
                dwl = dw.build_s1(g, scale=args.scale, keep_host=True)
                out = g.dalloc(total * rec_bytes(L))
                dw.synth_batch(g, dwl, out, total, 0, S1_SUB_RATE, S1_LOWQ_PROB, S1_LOWQ_CHARS, REC_ID_WIDTH)
                text = g.d2h(out, total * rec_bytes(L))
                index, name = dwl.host_index, dwl.name
    except Exception:
        index, text = None, None
    if text is None:
        wl = workloads.make_s1(scale=args.scale)
        index, name = wl.index, wl.name
        text = synth.simulate_reads(wl.genome, wl.haps, total, L, seed=wl.seed + 1000, sub_rate=S1_SUB_RATE, lowq_prob=S1_LOWQ_PROB,
                                    lowq_chars=S1_LOWQ_CHARS, id_width=REC_ID_WIDTH)
    bb = Bp * rec_bytes(L)

    if not use_ref:
        # the oracle port, one thread (the compiled reference is absent or would not fit in RAM)
        o = orc.Oracle(index)
        times = []
        for s in range(nb):
            t0 = time.perf_counter()
            o.process_fastq(text[s * bb:(s + 1) * bb], want_results=False)
            times.append(time.perf_counter() - t0)
        dt = sum(times[W:])
        value = K * Bp / dt
        kind, cores, sample = "port", 1, "oracle/liboracle.so, 1 thread, %d reads per step" % Bp
    else:
        d = tempfile.mkdtemp(prefix="vg_refarm_")
        prefix = os.path.join(d, "s1")
        ib.write_index(index, prefix)
        vcf = os.path.join(d, "snp.vcf")            # only opened after the read loop (src/qv.cc:1628), which this arm never reaches
        open(vcf, "w").write("##fileformat=VCFv4.0\n#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\n")
        procs, fifos, fds = [], [], []
        for p in range(nproc):
            ff = os.path.join(d, "reads%d.fq" % p)
            os.mkfifo(ff)
            fifos.append(ff)
            procs.append(subprocess.Popen([orc.REF_BIN, "geno", prefix, ff, vcf, os.path.join(d, "out%d.vcf" % p)],
                                          stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True))
        for p in range(nproc):
            fds.append(os.open(fifos[p], os.O_WRONLY))      # blocks until the process opens its FASTQ (after argv parsing)
        for p in range(nproc):                                # "Processing..." is printed once the index is loaded (src/qv.cc:753)
            for line in procs[p].stderr:
                if line.startswith("Processing"):
                    break

        def feed(p, s, res):
            t0 = time.perf_counter()
            off = (p * nb + s) * bb
            mv = memoryview(text)[off:off + bb]
            done = 0
            while done < bb:
                done += os.write(fds[p], mv[done:done + (1 << 20)])
            res[p] = time.perf_counter() - t0

        step_t = []
        for s in range(nb):
            res = [0.0] * nproc
            th = [threading.Thread(target=feed, args=(p, s, res)) for p in range(nproc)]
            t0 = time.perf_counter()
            for t in th:
                t.start()
            for t in th:
                t.join()
            step_t.append(time.perf_counter() - t0)
        for fd in fds:
            os.close(fd)
        for pr in procs:
            pr.kill()
        import shutil
        shutil.rmtree(d, ignore_errors=True)
        dt = sum(step_t[W:])
        value = K * Bp * nproc / dt
        kind, cores = "reference", nproc
        sample = ("unmodified reference `vargeno geno` (oracle/_ref), %d processes x 1 thread (the reference is single-threaded; "
                  "one process per core that fits in RAM at ~19 GB each), index loaded before timing, %d reads per process per step "
                  "pushed through a FIFO (time = writer completion, pipe slack 64 KiB)" % (nproc, Bp))
    out = {"impl": "reference", "metric": "reads/s", "value": value, "unit": "reads/s", "n_gpus": args.gpus, "steps": K, "warmup": W,
           "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
           "data": "synthetic",
           "config": {"workload": name + ", 1M-SNP list, 150bp reads @0.5% subst, lowq 0.25 on first 4 quality chars",
                      "reads_per_step": Bp * cores, "read_len": L},
           "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


_REAL_STDOUT = None


def emit(obj):
    """The one JSON line of this process, on the real stdout."""
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def main():
    # Libraries print to fd 1 on their own (NCCL's "NCCL version ..." banner at communicator set-up): everything but the
    # result line goes to stderr, so that stdout carries exactly one JSON line.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="shrink the genome / SNP list (tests only)")
    ap.add_argument("--workload", default="s1", choices=["s1", "s2", "s3"],
                    help="s1: chr22-shaped (BASELINE configs[1], the default); s2 / s3: GRCh38-shaped WGS / high-error stress (configs[2] / [3])")
    ap.add_argument("--batch-reads", type=int, default=2_000_000)
    ap.add_argument("--cpu-sample", type=int, default=400_000)
    ap.add_argument("--ref-batch-reads", type=int, default=100_000)
    ap.add_argument("--ref-procs", type=int, default=0)
    ap.add_argument("--clock-hold-s", type=float, default=1.0, help="extra seconds of the same load while nvidia-smi samples clocks")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-roofline-probe", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
