"""BGZF (blocked gzip, SAM/BAM specification section 4.1) writer and member scanner: test and bench tooling for the device-side
inflate (vgb_submit_bgzf).  A BGZF file is a series of gzip members, each with a 'BC' extra subfield that holds the member's
size, each inflating to at most 64 KiB; the file ends with an empty member."""
from __future__ import annotations

import struct
import zlib
from typing import List, Sequence, Tuple

import numpy as np

EOF_MEMBER = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def member(data: bytes, level: int = 6, strategy: int = zlib.Z_DEFAULT_STRATEGY) -> bytes:
    assert len(data) <= 65536
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    payload = c.compress(data) + c.flush()
    bsize = 12 + 6 + len(payload) + 8
    assert bsize <= 65536, "member too large: use smaller blocks for incompressible data"
    hdr = struct.pack("<BBBBIBBH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6) + b"BC" + struct.pack("<HH", 2, bsize - 1)
    return hdr + payload + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data))


def compress(data: bytes, block: int = 65280, level: int = 6, sizes: Sequence[int] = (), eof: bool = True) -> bytes:
    """data -> BGZF bytes.  block: uncompressed bytes per member (bgzip uses 65280); sizes: explicit member sizes, cycled
    (tests: tiny and uneven members)."""
    out, at, k = [], 0, 0
    while at < len(data):
        n = sizes[k % len(sizes)] if sizes else block
        k += 1
        out.append(member(data[at:at + n], level))
        at += n
    if eof:
        out.append(EOF_MEMBER)
    return b"".join(out)


def scan(buf: bytes) -> List[Tuple[int, int, int, int]]:
    """[(member offset, member size, payload offset inside the member, ISIZE)] -- the walk the C++ host does (geno_host.cpp bgzf_scan)."""
    out, off = [], 0
    while off < len(buf):
        assert buf[off:off + 4] == b"\x1f\x8b\x08\x04", "not a BGZF member at %d" % off
        xlen = struct.unpack_from("<H", buf, off + 10)[0]
        x, bsize = 0, 0
        while x + 4 <= xlen:
            si1, si2, slen = buf[off + 12 + x], buf[off + 13 + x], struct.unpack_from("<H", buf, off + 14 + x)[0]
            if si1 == 66 and si2 == 67 and slen == 2:
                bsize = struct.unpack_from("<H", buf, off + 16 + x)[0] + 1
            x += 4 + slen
        assert bsize
        isize = struct.unpack_from("<I", buf, off + bsize - 4)[0]
        out.append((off, bsize, 12 + xlen, isize))
        off += bsize
    return out


def plan_chunks(members, own_out: int, overlap: int = 8192):
    """[(first overlap member, first own member, end member, overlap bytes)] over the non-empty members: what stream_bgzf plans."""
    mem = [m for m in members if m[3]]
    plan, i = [], 0
    while i < len(mem):
        i0, out = i, 0
        while i < len(mem) and (out + mem[i][3] <= own_out or i == i0):
            out += mem[i][3]
            i += 1
        o0, ov = i0, 0
        while o0 > 0 and ov < overlap:
            o0 -= 1
            ov += mem[o0][3]
        plan.append((o0, i0, i, ov))
    return mem, plan


def chunk_arrays(buf: bytes, mem, o0: int, i1: int):
    """compressed bytes of members [o0, i1) back to back + their (payload offset, payload length, ISIZE) table for vgb_submit_bgzf"""
    parts, tab, at = [], [], 0
    for off, size, poff, isize in mem[o0:i1]:
        parts.append(buf[off:off + size])
        tab.append((at + poff, size - poff - 8, isize))
        at += size
    return np.frombuffer(b"".join(parts), dtype=np.uint8), np.array(tab, dtype="<u4").reshape(-1, 3)


def _span(args):
    path, start, n, block, level = args
    with open(path, "rb") as f:
        f.seek(start)
        data = f.read(n)
    return b"".join(member(data[a:a + block], level) for a in range(0, len(data), block))


def compress_file(src: str, dst: str, procs: int = 8, block: int = 65280, level: int = 1, span_blocks: int = 512) -> None:
    """src -> BGZF dst with `procs` worker processes (what `bgzip -@ N` does); spans are whole numbers of members."""
    import multiprocessing as mp
    import os
    size = os.path.getsize(src)
    span = block * span_blocks
    jobs = [(src, a, min(span, size - a), block, level) for a in range(0, size, span)]
    with mp.Pool(procs) as pool, open(dst, "wb") as out:
        for piece in pool.imap(_span, jobs, chunksize=1):
            out.write(piece)
        out.write(EOF_MEMBER)


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="write a file as BGZF (multi-process)")
    ap.add_argument("src")
    ap.add_argument("dst")
    ap.add_argument("--procs", type=int, default=8)
    ap.add_argument("--level", type=int, default=1)
    a = ap.parse_args()
    compress_file(a.src, a.dst, a.procs, level=a.level)
