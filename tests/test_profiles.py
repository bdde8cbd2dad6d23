"""profiles/ncu_constants.json holds per-launch ncu counters of the dominant kernel (roofline.traffic and the request-rate keys of
bench.py).  They are tied to a hash of the kernel sources and bench.py drops them when the hash differs: a kernel edit without a
fresh capture must show up here, not as a silently missing `traffic` in the next benchmark line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_ncu_constants_belong_to_the_current_kernel_sources():
    import bench
    c = json.load(open(os.path.join(ROOT, "profiles", "ncu_constants.json")))
    assert {"s2", "s3"} <= set(c)
    for wl in ("s2", "s3"):
        assert c[wl]["kernel_hash"] == bench.kernel_hash(), "%s: recapture with `ncu --set full` and tools/ncu_constants.py" % wl
        assert os.path.exists(os.path.join(ROOT, c[wl]["source"]))
        assert c[wl]["reads_per_launch"] == 2_000_000 and c[wl]["dram_bytes_per_read"] > 0 and c[wl]["l2_requests_per_read"] > 0
    nc = bench.ncu_constants("s2")
    assert nc is not None and nc["stale"] is False
