"""world_size-2 run of the multi-GPU host logic on CPU (gloo): round-robin chunk dealing + one counter all-reduce.
Each rank stands in for a GPU with the CPU oracle as its worker; the reduced, saturated counters must equal a
single-process run over all reads."""
import os
import tempfile

import numpy as np
import pytest

from vargeno_b200 import sharding


def _worker(rank, world, init_file, ds_dir, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    from oracle import oracle as orc
    from vargeno_b200.geno import Genotyper
    from vargeno_b200.tools import index_builder as ib

    dist.init_process_group("gloo", init_method="file://" + init_file, rank=rank, world_size=world)
    ix = ib.build_index(os.path.join(ds_dir, "ref.fa"), os.path.join(ds_dir, "snp.vcf"))
    fq = np.fromfile(os.path.join(ds_dir, "reads.fq"), dtype=np.uint8)
    chunks = Genotyper.split_records(fq, 256 << 10)
    mine = sharding.deal_chunks(len(chunks), rank, world)
    o = orc.Oracle(ix)
    for i in mine:
        s, e, _ = chunks[i]
        o.process_fastq(fq[s:e], want_results=False)
    sites = o.sites()
    # unsaturated per-rank counters are what the GPUs hold; the oracle saturates per rank, which is legal for the
    # sum-then-clamp reduction as long as the reduce type is wide enough
    cnt = torch.from_numpy(np.stack([sites["ref_cnt"], sites["alt_cnt"]], axis=1).astype(np.int64))
    sharding.allreduce_counts(cnt)
    cnt = sharding.saturate(cnt)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), cnt.numpy())
        np.save(os.path.join(out_dir, "n_chunks.npy"), np.array([len(chunks)]))
    dist.destroy_process_group()


def test_two_rank_reduction_equals_single_run(cache):
    import torch.multiprocessing as mp
    from oracle import oracle as orc
    ds = cache.dataset("s0")
    with tempfile.TemporaryDirectory() as d:
        init = os.path.join(d, "init")
        mp.spawn(_worker, args=(2, init, ds.dir, d), nprocs=2, join=True)
        reduced = np.load(os.path.join(d, "reduced.npy"))
        assert int(np.load(os.path.join(d, "n_chunks.npy"))[0]) > 4
    o = orc.Oracle(cache.index("s0"))
    o.process_fastq(np.fromfile(ds.fastq, dtype=np.uint8), want_results=False)
    s = o.sites()
    assert np.array_equal(reduced[:, 0], s["ref_cnt"]) and np.array_equal(reduced[:, 1], s["alt_cnt"])
    o.close()


def test_dealing_covers_every_chunk_once():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in sharding.deal_chunks(29, r, world))
        assert seen == list(range(29))
    ids = [sharding.shard_read_ids([5, 7, 3, 9], r, 2) for r in range(2)]
    assert ids[0] == [(0, 0, 5), (2, 12, 3)] and ids[1] == [(1, 5, 7), (3, 15, 9)]
    with pytest.raises(ValueError):
        sharding.deal_chunks(4, 2, 2)
