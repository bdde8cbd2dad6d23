"""Host-side logic of the multi-GPU path: reads are independent units, the index is replicated,
record-aligned FASTQ chunks are dealt round robin to the ranks, and the only exchange is one sum of the per-SNP
{ref_cnt, alt_cnt} counters before calling (saturation at 63 is applied after the sum: min(63, sum_g c_g)).

The C++ host (csrc/host/geno_host.cpp) and bench.py use NCCL inside libvgb200.so (vgb_allreduce_pileup); the helpers
here express the same dealing / reduction with torch.distributed so that they can be exercised with the gloo backend on
CPU-only machines (tests/test_sharding_gloo.py) and used with an external process group on raw device counters
(vgb_counter_device_ptr)."""
from __future__ import annotations

from typing import List, Sequence, Tuple

MAX_COV = 63   # src/vartype.h:27


def deal_chunks(n_chunks: int, rank: int, world: int) -> List[int]:
    """Chunk i goes to rank i % world (the order stream_fastq in csrc/host/geno_host.cpp uses)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_chunks, world))


def shard_read_ids(chunk_reads: Sequence[int], rank: int, world: int) -> List[Tuple[int, int, int]]:
    """(chunk index, first global read id, n reads) of the chunks this rank owns."""
    first, out = 0, []
    for i, n in enumerate(chunk_reads):
        if i % world == rank:
            out.append((i, first, n))
        first += n
    return out


def allreduce_counts(counts, group=None):
    """In-place sum over ranks of an integer tensor of unsaturated counters (CPU tensor with gloo, CUDA tensor with nccl)."""
    import torch.distributed as dist
    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


def saturate(counts):
    return counts.clamp(max=MAX_COV)
