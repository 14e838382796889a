#!/usr/bin/env python
"""Generates tests/golden/ikd_*.npz from the REFERENCE's own ikd-Tree (oracle/_ref/libref_ikd.so, built from
/root/reference/include/ikd-Tree by oracle/Makefile.ref).  Run in the authoring container (needs /root/reference):

    python tools/make_golden.py

Fixtures (inputs + reference outputs), consumed by tests/test_ref_ikd.py (oracle, CPU) and tests/test_mesh_gpu.py (CUDA):
  ikd_knn.npz     vertices (insertion order), queries, Nearest_Search results for (k, max_dist) in CASES
  ikd_append.npz  three world-frame frames (sub-sampled) and the vertex list produced by the reference's append loop
                  (pointcloud_rgbd.cpp:411-552: xi-grid test, ikd-Tree 1-NN test, Add_Point) driven with the real tree
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_ikd  # noqa: E402
from immesh_b200 import api, synth  # noqa: E402

CASES = [(20, float("inf")), (1, 0.1), (20, 1.0), (5, 0.3), (32, 0.75)]


def append_with_tree(frames, xi, target, tree=None):
    """Global_map::append_points_to_global_map (pointcloud_rgbd.cpp:411-552) with the kd-tree queries answered by `tree`."""
    tree = tree or ref_ikd.RefIkdTree()
    verts, grid = [], set()
    for pts in frames:
        step = max(1, round(len(pts) // target))
        for i in range(0, len(pts), step):
            p = pts[i]
            g = tuple(int(np.round(np.float64(p[j]) / xi)) for j in range(3))
            if g in grid:
                continue
            if verts:
                idx, d2, cnt = tree.knn(p[None, :], 1)
                if cnt[0] and np.sqrt(d2[0, 0]) < xi:
                    continue
            grid.add(g)
            verts.append(p.copy())
            tree.add(p[None, :])
    return np.asarray(verts, dtype=np.float32)


def main():
    out = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out, exist_ok=True)
    cfg = api.MeshConfig()
    sensor, scans = synth.make_stream("avia", 3, seed=21, n_points=5000)
    frames = [(s["body_full"].astype(np.float64) @ s["R_true"].T + s["t_true"]).astype(np.float32) for s in scans]
    tree = ref_ikd.RefIkdTree()
    verts = append_with_tree(frames, cfg.points_minimum_scale, cfg.number_of_pts_append_to_map, tree)
    np.savez_compressed(os.path.join(out, "ikd_append.npz"), **{f"frame{k}": f for k, f in enumerate(frames)},
                        pose_t=np.stack([s["t_true"] for s in scans]), verts=verts, xi=cfg.points_minimum_scale,
                        res=cfg.voxel_resolution, target=cfg.number_of_pts_append_to_map)
    rng = np.random.default_rng(3)
    q = (verts[rng.integers(0, len(verts), 192)] + rng.normal(0, 0.15, (192, 3))).astype(np.float32)
    q = np.concatenate([q, verts[:32], (verts[:32] + np.float32(50.0))]).astype(np.float32)   # exact hits and far misses
    d = {"verts": verts, "queries": q, "cases": np.array(CASES)}
    for c, (k, md) in enumerate(CASES):
        idx, d2, cnt = tree.knn(q, k, md)
        d[f"idx{c}"], d[f"d2{c}"], d[f"cnt{c}"] = idx.astype(np.int32), d2, cnt
    np.savez_compressed(os.path.join(out, "ikd_knn.npz"), **d)
    print("wrote", os.listdir(out), "vertices:", len(verts))


if __name__ == "__main__":
    main()
