#!/usr/bin/env python
"""The localization half of a bench workload with the blocking API (no mesher): target for `ncu --cache-control none` launch lists of
the IESKF kernels on a warm map.
    python tools/debug/lio_only.py C100k 30
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402
from immesh_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C100k"
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 20
wl = bench.workloads()[name]
scans = bench.get_stream(wl, n_scans)
g = api.Lio(wl["lio"])
g.set_state(bench.init_state_vec(scans))
g.voxel_map_init(scans[0]["body_full"])
for k in range(1, n_scans):
    s, it = g.step(scans[k]["body_ds"], scans[k]["dt"])
print("ok", name, n_scans, "iters last", it, g.counts())
