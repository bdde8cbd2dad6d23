"""CPU-side checks of the boundary: the shared library loads without a GPU and exports exactly what
include/vgb200.h declares; compute entry points fail loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vargeno_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_and_library_agree(lib):
    from vargeno_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "vgb200.h")).read()
    declared = set(re.findall(r"\b(vgb_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vgb_abi_version() == 2


def test_struct_layouts_match_header():
    from vargeno_b200 import _lib
    assert C.sizeof(_lib.Config) == 32 and C.sizeof(_lib.IndexView) == 14 * 8
    assert _lib.HIT.itemsize == 32 and _lib.READ_RESULT.itemsize == 24


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from vargeno_b200.geno import Genotyper, VgbError
    with pytest.raises(VgbError) as ei:
        Genotyper(device=0)
    assert ei.value.code == -2 and "no CPU fallback" in str(ei.value)


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or execute anything under oracle/."""
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "vargeno_b200")):
        if "_obj" in base or "__pycache__" in base:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|liboracle|vg_oracle\.h|oracle/_ref", text):
                    bad.append(os.path.join(base, f))
    assert not bad, bad


def test_record_splitter():
    from vargeno_b200.geno import Genotyper
    recs = [b"@r%d\n%s\n+\n%s\n" % (i, b"ACGT" * (3 + i % 5), b"I" * (4 * (3 + i % 5))) for i in range(50)]
    text = np.frombuffer(b"".join(recs), dtype=np.uint8)
    chunks = Genotyper.split_records(text, 200)
    assert chunks[0][0] == 0 and chunks[-1][1] == text.size and sum(c[2] for c in chunks) == 50
    for (s, e, n), nxt in zip(chunks, chunks[1:] + [None]):
        assert e - s <= 200 and np.count_nonzero(text[s:e] == 10) == 4 * n
        if nxt:
            assert nxt[0] == e
