import os
import sys
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box: pytest -m gpu)")
    config.addinivalue_line("markers", "ref: re-runs the compiled reference (oracle/_ref; ~16 GiB RAM, ~1 min per data set)")


def pytest_collection_modifyitems(config, items):
    # the `ref` tests are opt-in: they need oracle/_ref and a lot of memory
    if os.environ.get("VG_RUN_REF") == "1":
        return
    skip = pytest.mark.skip(reason="set VG_RUN_REF=1 to re-run the compiled reference")
    for it in items:
        if "ref" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def data_root():
    with tempfile.TemporaryDirectory(prefix="vg_tests_") as d:
        yield d


class _Cache:
    """Builds each named data set (and its index) once per session."""

    def __init__(self, root):
        self.root = root
        self.ds = {}
        self.ix = {}

    def dataset(self, name):
        import datasets
        if name not in self.ds:
            self.ds[name] = datasets.MAKERS[name](os.path.join(self.root, name))
        return self.ds[name]

    def index(self, name):
        from vargeno_b200.tools import index_builder as ib
        if name not in self.ix:
            ds = self.dataset(name)
            self.ix[name] = ib.build_index(ds.fasta, ds.vcf)
        return self.ix[name]


@pytest.fixture(scope="session")
def cache(data_root):
    return _Cache(data_root)
