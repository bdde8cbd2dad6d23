"""Host logic of the FASTQ feed (csrc/host/geno_host.cpp): chunks handed to the GPU must start and end at record boundaries
(4 lines), lose no byte, and carry the right first-read ordinal -- for plain, gzip and several files read back to back, at chunk
sizes from "barely one record" upwards.  No GPU: `vargeno-b200 fastq-chunks` runs the same chunker against a printing sink."""
import gzip
import subprocess

import numpy as np
import pytest

from vargeno_b200 import build as vb


def _fnv(b):
    h = 1469598103934665603
    for x in b:
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def _fastq(n, seed, trailing_newline=True):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        L = int(rng.integers(31, 260))
        seq = "".join("ACGTN"[j] for j in rng.integers(0, 5, L))
        qual = "".join(chr(33 + int(q)) for q in rng.integers(0, 41, L))
        qual = "@" + qual[1:] if i % 7 == 0 else qual        # quality lines that start with '@' (the classic FASTQ ambiguity)
        recs.append("@r%d some description\n%s\n+\n%s\n" % (i * 977, seq, qual))
    text = "".join(recs)
    return (text if trailing_newline else text[:-1]).encode()


def _run(files, chunk_bytes):
    vb.build()
    p = subprocess.run([vb.HOST_BIN, "fastq-chunks", files, "--chunk-bytes", str(chunk_bytes)], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    return p


@pytest.mark.parametrize("chunk_bytes", [700, 4096, 100_000, 1 << 22])
@pytest.mark.parametrize("layout", ["plain", "gz", "split_mid_record", "no_trailing_newline"])
def test_chunks_are_record_aligned_and_lossless(tmp_path, chunk_bytes, layout):
    text = _fastq(400, seed=3, trailing_newline=layout != "no_trailing_newline")
    if layout == "gz":
        f = tmp_path / "a.fq.gz"
        with gzip.open(f, "wb") as g:
            g.write(text)
        files = str(f)
    elif layout == "split_mid_record":
        cut = len(text) // 3 + 5                                   # inside a record: the mates-back-to-back case, worst alignment
        a, b, c = tmp_path / "a.fq", tmp_path / "b.fq.gz", tmp_path / "c.fq"
        a.write_bytes(text[:cut])
        with gzip.open(b, "wb") as g:
            g.write(text[cut:2 * cut])
        c.write_bytes(text[2 * cut:])
        files = "%s,%s,%s" % (a, b, c)
    else:
        f = tmp_path / "a.fq"
        f.write_bytes(text)
        files = str(f)
    p = _run(files, chunk_bytes)
    assert p.returncode == 0, p.stderr
    lines = p.stdout.strip().splitlines()
    assert lines[-1].startswith("total %d bytes" % len(text))
    off, reads = 0, 0
    for k, ln in enumerate(lines[:-1]):
        idx, nbytes, nl, first, h = ln.split()
        nbytes, nl, first = int(nbytes), int(nl), int(first)
        chunk = text[off:off + nbytes]
        assert int(idx) == k and nbytes <= chunk_bytes and first == reads
        assert _fnv(chunk) == int(h, 16), "chunk %d does not continue where the previous one ended" % k
        assert chunk.count(b"\n") == nl
        last = k == len(lines) - 2
        n_lines = nl + (1 if last and not chunk.endswith(b"\n") else 0)
        assert n_lines % 4 == 0 and (last or chunk.endswith(b"\n")), "chunk %d is not a whole number of records" % k
        assert chunk[:1] == b"@"
        off += nbytes
        reads += n_lines // 4
    assert off == len(text) and reads == 400


def test_record_larger_than_chunk_is_an_error(tmp_path):
    f = tmp_path / "a.fq"
    f.write_bytes(_fastq(3, seed=1))
    p = _run(str(f), 64)
    assert p.returncode != 0 and "larger than the chunk size" in p.stderr


def test_missing_file_is_an_error(tmp_path):
    p = _run(str(tmp_path / "nope.fq"), 4096)
    assert p.returncode != 0 and "cannot open" in p.stderr


def test_parallel_pread_gives_the_same_chunks(tmp_path):
    """Large reads of a plain file are split over several pread()s running in parallel (VGB_READ_THREADS): same chunks, byte for
    byte, as the single-threaded reader."""
    import os
    vb.build()
    block = _fastq(2000, seed=9)
    f = tmp_path / "big.fq"
    with open(f, "wb") as out:
        for _ in range(60):
            out.write(block)                    # ~45 MB
    outs = []
    for t in ("1", "5"):
        env = dict(os.environ, VGB_READ_THREADS=t)
        p = subprocess.run([vb.HOST_BIN, "fastq-chunks", str(f), "--chunk-bytes", str(20 << 20)], stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, text=True, env=env)
        assert p.returncode == 0, p.stderr
        outs.append(p.stdout)
    assert outs[0] == outs[1]
    assert outs[0].strip().splitlines()[-1].startswith("total %d bytes" % (60 * len(block)))


@pytest.mark.parametrize("feeders", [1, 3, 8])
@pytest.mark.parametrize("chunk_bytes", [300_000, 1 << 20, 1 << 23])
@pytest.mark.parametrize("layout", ["one_file", "three_files_cut_mid_record", "no_trailing_newline"])
def test_parallel_reader_tiles_the_input_with_whole_records(tmp_path, feeders, chunk_bytes, layout):
    """The per-GPU feeders of plain files (stream_plain_parallel): segments are claimed independently, record starts are found
    locally ('@' at a line start whose second next line starts with '+': every seventh quality line here starts with '@'), and
    the chunks must tile the input exactly -- whatever the segment size, the number of feeders and the file boundaries."""
    vb.build()
    text = _fastq(6000, seed=9, trailing_newline=layout != "no_trailing_newline")
    if layout == "three_files_cut_mid_record":
        cut = len(text) // 3 + 11
        parts = [text[:cut], text[cut:2 * cut], text[2 * cut:]]
    else:
        parts = [text]
    names = []
    for i, part in enumerate(parts):
        f = tmp_path / ("p%d.fq" % i)
        f.write_bytes(part)
        names.append(str(f))
    p = subprocess.run([vb.HOST_BIN, "fastq-chunks", ",".join(names), "--chunk-bytes", str(chunk_bytes), "--parallel", str(feeders)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert p.returncode == 0, p.stderr
    rows = sorted([int(x, 16) if k == 3 else int(x) for k, x in enumerate(l.split())] for l in p.stdout.splitlines() if not l.startswith("total"))
    assert len(rows) >= 1
    at = 0
    for off, nbytes, lines, h in rows:
        assert off == at, "gap or overlap at byte %d" % at
        piece = text[off:off + nbytes]
        assert piece[:1] == b"@" and _fnv(piece) == h
        n_lines = piece.count(b"\n") + (0 if piece.endswith(b"\n") else 1)
        assert n_lines % 4 == 0                                    # whole records only
        at += nbytes
    assert at == len(text)
    assert len(rows) >= min(3, len(text) // chunk_bytes)
