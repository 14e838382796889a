"""Work distribution of the match walk (build_single_residual) over the points of a scan, from the host emulation of the kernels:
how many points sit in a split root voxel, how many planes their walk evaluates and how many of those pass the range gate.
    python tools/debug/match_stats.py C100k 10
"""
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402
from immesh_b200 import api, build  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C100k"
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 8
wl = bench.workloads()[name]
lib = api.load_library(build.build_emu())
scans = bench.get_stream(wl, n_scans)
g = api.Lio(wl["lio"], lib=lib)
g.set_state(bench.init_state_vec(scans))
g.voxel_map_init(scans[0]["body_full"])
for k in range(1, n_scans):
    body = scans[k]["body_ds"]
    g.predict(scans[k]["dt"])
    g.lio_state_estimation(body)
    n = body.shape[0]
    out = np.zeros((n, 4), np.int32)
    lib.emu_match_walk_stats(g._h, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n)
    kind, nodes, planes, heavy = out.T
    sp = kind == 2
    w = heavy.reshape(-1)[: n // 32 * 32].reshape(-1, 32)
    print(f"scan {k}: n {n}  no-voxel {np.mean(kind == 0):.3f}  plane-root {np.mean(kind == 1):.3f}  split-root {np.mean(sp):.3f} | split roots: "
          f"nodes mean {nodes[sp].mean():.1f} max {nodes[sp].max()}  planes mean {planes[sp].mean():.1f} max {planes[sp].max()}  "
          f"gate-pass mean {heavy[sp].mean():.2f} max {heavy[sp].max()} | per warp of 32 consecutive points: max gate-pass mean {w.max(1).mean():.2f}, sum mean {w.sum(1).mean():.1f}")
    lo = np.zeros((n, 8, 3), np.int32)
    lib.emu_match_lane_stats(g._h, lo.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n)
    w4 = lo[: n // 4 * 4].reshape(-1, 32, 3)     # one warp of k_match: 4 points x 8 lanes
    print(f"        8 lanes per point, per warp (4 points): slowest lane visits {w4[:, :, 0].max(1).mean():.1f} nodes (max {w4[:, :, 0].max()}), "
          f"evaluates {w4[:, :, 1].max(1).mean():.1f} planes (max {w4[:, :, 1].max()}), {w4[:, :, 2].max(1).mean():.2f} past the gate (max {w4[:, :, 2].max()})")
    g.map_incremental_grow()
