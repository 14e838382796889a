#!/usr/bin/env python
"""Latency budget of the per-voxel triangulation (k_voxel_tri_warp) from the in-kernel clock64 stamps of the debug library
(tools/debug/build_stamps.sh): warp 0 of block 0, the LAST voxel it processed in the last frame.
    python tools/debug/mesh_stamps.py C100k 16"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402
from immesh_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C100k"
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 16
lib = api.load_library(os.path.join(ROOT, "tools", "debug", "libimmesh_stamps.so"))
wl = bench.workloads()[name]
scans = bench.get_stream(wl, n_scans)
g = api.Lio(wl["lio"], lib=lib)
mesh = api.Mesh(wl["mesh"], lib=lib)
g.set_state(bench.init_state_vec(scans))
g.voxel_map_init(scans[0]["body_full"])
for k in range(1, n_scans):
    g.step(scans[k]["body_ds"], scans[k]["dt"])
    mesh.push_frame_from_lio(g, scans[k]["body_full"])
    st = (C.c_longlong * 64)()
    assert lib.immesh_debug_stamps_mesh(st) == 0
    s = list(st)
    d = lambda a, b: s[b] - s[a]
    if k >= n_scans - 6:
        print(f"frame {k}: warp 0 of block 0: kernel {d(30, 31)} cycles for {s[41]} voxels of {s[42]} | last voxel: n {s[39]} faces {s[40]} total {d(32, 38)} = "
              f"load {d(32, 33)}  PCA(lane 0) {d(33, 34)}  project+seed {d(34, 35)}  Bowyer-Watson {d(35, 36)} ({d(35, 36) // max(1, s[39] - 3)}/insertion)  "
              f"faces {d(36, 37)}  output {d(37, 38)}")
        print(f"          PCA: centroid + covariance {d(33, 51)}  Jacobi {d(51, 52)}  ordering + axes {d(52, 34)}")
        print(f"          k_cand_init, candidate 0: point load {d(53, 54)}  voxel get-or-create {d(54, 55)}  activation {d(55, 56)}  xi-cell lookup {d(56, 57)}  "
              f"27 probes + vertex ids {d(57, 58)}  vertex positions {d(58, 59)}  frame grid insert + list push {d(59, 60)}  (stale when the candidate left early)")
        ins = max(1, s[39] - 3)
        print(f"          per insertion: prefilter scan {s[44] // ins}  exact conflicts {s[45] // ins}  cavity edges + new triangles {s[46] // ins}  tail {s[47] // ins} | "
              f"pool {s[50] / ins:.1f} triangles, {s[48] / ins:.1f} past the prefilter, {s[49] / ins:.1f} in conflict")
