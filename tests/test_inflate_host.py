"""The device DEFLATE decoder (vargeno_b200/csrc/vgb_inflate.cuh) compiled for the HOST with one lane, against zlib: every block
type (stored, fixed, dynamic), every compression level, long codes (beyond the 10-bit direct table), overlapping matches,
FASTQ-like text, empty input, and corrupted / truncated streams that must be reported, not decoded.  The GPU runs the same
source with 32 lanes (tests/test_gpu_bgzf.py)."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("inflate") / "libinflate_host.so")
    subprocess.check_call(["g++", "-O2", "-Wall", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "native", "inflate_host.cpp")])
    L = C.CDLL(so)
    L.vgb_host_inflate.argtypes = [C.c_char_p, C.c_ulonglong, C.c_char_p, C.c_uint, C.POINTER(C.c_uint)]
    return L


def _inflate(lib, comp, cap):
    out = C.create_string_buffer(max(cap, 1))
    n = C.c_uint()
    rc = lib.vgb_host_inflate(comp, len(comp), out, cap, C.byref(n))
    return rc, out.raw[:n.value]


def _deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, 9, strategy)
    return c.compress(data) + c.flush()


def _fastq(n, seed):
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        seq = "".join("ACGT"[j] for j in rng.integers(0, 4, 150))
        qual = "".join(chr(35 + int(q)) for q in np.minimum(40, rng.geometric(0.08, 150)))
        recs.append("@r%09d\n%s\n+\n%s\n" % (i, seq, qual))
    return "".join(recs).encode()


def _rnd(n, seed):
    return np.random.default_rng(seed).integers(0, 256, n, dtype=np.uint8).tobytes()


CASES = {
    "empty": b"",
    "one_byte": b"A",
    "fastq": _fastq(200, 1),                                            # 64 KB, like a BGZF block of reads
    "zeros": bytes(65536),                                               # distance-1 matches of length 258
    "period3": b"ACG" * 20000,                                           # overlapping matches, distance < length
    "random": np.random.default_rng(2).integers(0, 256, 65536, dtype=np.uint8).tobytes(),     # incompressible: stored blocks
    "skewed": bytes(np.minimum(255, np.random.default_rng(3).geometric(0.02, 60000)).astype(np.uint8)),   # long codes (> 10 bits)
    "text": (b"the quick brown fox jumps over the lazy dog. " * 1400)[:65000],
    # matches that reach further back than the decoder's 8 KiB ring (read back from the flushed output), into a stored block that
    # is still in the ring, and into one that is not
    "far_repeat": _rnd(20000, 4) * 3,
    "stored_then_near": _rnd(3000, 5) * 4 + b"ACGT" * 500 + _rnd(3000, 5),
    "stored_then_far": _rnd(30000, 6) + _rnd(30000, 6)[:12000] + b"N" * 700 + _rnd(30000, 6)[5000:9000],
}


@pytest.mark.parametrize("level", [0, 1, 6, 9])
@pytest.mark.parametrize("name", sorted(CASES))
def test_matches_zlib(lib, name, level):
    data = CASES[name]
    comp = _deflate(data, level)
    rc, out = _inflate(lib, comp, len(data))
    assert rc == 0 and out == data


@pytest.mark.parametrize("name", ["fastq", "period3", "text", "empty"])
def test_fixed_huffman_blocks(lib, name):
    data = CASES[name]
    comp = _deflate(data, 6, zlib.Z_FIXED)
    rc, out = _inflate(lib, comp, len(data))
    assert rc == 0 and out == data


def test_huffman_only_and_rle_strategies(lib):
    for strat in (zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FILTERED):
        for name in ("fastq", "skewed"):
            data = CASES[name]
            rc, out = _inflate(lib, _deflate(data, 6, strat), len(data))
            assert rc == 0 and out == data


def test_several_blocks_in_one_stream(lib):
    """Z_FULL_FLUSH ends a block (and emits an empty stored block): block types change inside one stream."""
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    data = CASES["fastq"][:20000] + CASES["random"][:5000] + CASES["zeros"][:9000]
    comp = c.compress(data[:20000]) + c.flush(zlib.Z_FULL_FLUSH) + c.compress(data[20000:25000]) + c.flush(zlib.Z_SYNC_FLUSH) + \
        c.compress(data[25000:]) + c.flush()
    rc, out = _inflate(lib, comp, len(data))
    assert rc == 0 and out == data


def test_errors_are_reported(lib):
    data = CASES["fastq"]
    comp = _deflate(data, 6)
    rc, _ = _inflate(lib, comp, len(data) - 1)                          # output buffer one byte short
    assert rc != 0
    rc, out = _inflate(lib, comp[:len(comp) // 2], len(data))           # truncated input
    assert rc != 0 or out != data
    bad = bytearray(comp)
    bad[0] |= 0x06                                                       # BTYPE = 3 (reserved)
    rc, _ = _inflate(lib, bytes(bad), len(data))
    assert rc != 0
    rc, _ = _inflate(lib, b"\x01\x05\x00\x00\x00hello", 5)               # stored block whose NLEN is not ~LEN
    assert rc != 0
    rng = np.random.default_rng(5)
    for _ in range(200):                                                  # random garbage never crashes or overruns
        junk = rng.integers(0, 256, int(rng.integers(1, 400)), dtype=np.uint8).tobytes()
        rc, out = _inflate(lib, junk, 4096)
        assert len(out) <= 4096


@pytest.mark.parametrize("n_stored", [100, 3000, 7900, 20000, 65535])
def test_matches_into_a_stored_block(lib, n_stored):
    """stored block + a compressed block whose matches point into it (written by zlib against a preset dictionary): near ones
    come out of the decoder's ring, far ones out of the text that is already flushed."""
    first = _rnd(n_stored, 7)
    second = first[-5000:] + b"ACGT" * 300 + first[:4000] + first[n_stored // 2:n_stored // 2 + 300]
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 9, zlib.Z_DEFAULT_STRATEGY, first[-32768:])
    tail = c.compress(second) + c.flush()
    stored = bytes([0]) + len(first).to_bytes(2, "little") + (len(first) ^ 0xFFFF).to_bytes(2, "little") + first
    rc, out = _inflate(lib, stored + tail, len(first) + len(second))
    assert rc == 0 and out == first + second
