"""CPU tier: the committed ncu figures bench.py reads for `roofline` (profiles/traffic_r2.json) were captured on the
kernel sources in this tree -- otherwise every bench line would carry traffic_stale: true."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_committed_capture_matches_the_kernel_sources():
    import bench
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic_r2.json")))
    assert d["kernel_source_hash"] == bench.kernel_source_hash(), (
        "kernel sources changed after the capture: re-run tools/capture_profiles.sh k1g + profiles/make_capture.py")
    cap, stale = bench.committed_capture()
    assert cap is not None and not stale
    k = d["selfplay_k1g_kernel"]
    assert k["warp_inst_per_sim"] > 0 and k["dram_bytes_per_step"] > 0
