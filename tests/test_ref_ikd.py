"""Pins the kNN / vertex-append rows of the oracle to the REFERENCE's own ikd-Tree.

* live: oracle vs oracle/_ref/libref_ikd.so (include/ikd-Tree/ikd_Tree.cpp compiled unmodified, oracle/Makefile.ref) --
  skipped where the library is neither built nor buildable;
* golden: oracle vs tests/golden/ikd_*.npz, generated from the same library by tools/make_golden.py (always runs).
The CUDA path is checked against the same fixtures in tests/test_mesh_gpu.py."""
import os

import numpy as np
import pytest

import oracle_api as oa
import ref_ikd
from immesh_b200 import api, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
needs_ref = pytest.mark.skipif(not ref_ikd.available(), reason="oracle/_ref/libref_ikd.so not built and /root/reference absent")


def _oracle_with_frames(frames, pose_t, cfg):
    o = oa.OracleMesh(cfg)
    for k, f in enumerate(frames):
        o.push_frame(f, pose_t[k], k)
    return o


def test_golden_append_matches_oracle():
    g = np.load(os.path.join(GOLD, "ikd_append.npz"))
    cfg = api.MeshConfig()
    assert float(g["xi"]) == cfg.points_minimum_scale and float(g["res"]) == cfg.voxel_resolution
    frames = [g[f"frame{k}"] for k in range(3)]
    o = _oracle_with_frames(frames, g["pose_t"], cfg)
    v, _, _ = o.snapshot()
    assert v.shape == g["verts"].shape and np.array_equal(v, g["verts"])        # ids = append order, positions bit-exact


def test_golden_knn_matches_oracle():
    g = np.load(os.path.join(GOLD, "ikd_knn.npz"))
    ga = np.load(os.path.join(GOLD, "ikd_append.npz"))
    o = _oracle_with_frames([ga[f"frame{k}"] for k in range(3)], ga["pose_t"], api.MeshConfig())
    for c, (k, md) in enumerate(g["cases"]):
        idx, d2 = o.knn(g["queries"], int(k), float(md))
        assert ref_ikd.same_knn(idx.astype(np.int64), d2, g[f"idx{c}"].astype(np.int64), g[f"d2{c}"]), (k, md)


@needs_ref
def test_oracle_knn_matches_reference_ikdtree_incremental():
    sensor, scans = synth.make_stream("avia", 4, seed=11, n_points=12000)
    frames = [(s["body_full"].astype(np.float64) @ s["R_true"].T + s["t_true"]).astype(np.float32) for s in scans]
    o, t, n_prev = oa.OracleMesh(api.MeshConfig()), ref_ikd.RefIkdTree(), 0
    rng = np.random.default_rng(5)
    for k, f in enumerate(frames):
        o.push_frame(f, scans[k]["t_true"], k)
        v, _, _ = o.snapshot()
        t.add(v[n_prev:])            # Add_Point per accepted vertex, as pointcloud_rgbd.cpp:540 (tree re-balances itself)
        n_prev = len(v)
        q = (v[rng.integers(0, len(v), 1500)] + rng.normal(0, 0.15, (1500, 3))).astype(np.float32)
        for kk, md in ((20, float("inf")), (1, 0.1), (20, 1.0), (5, 0.3)):
            io, do = o.knn(q, kk, md)
            ir, dr, _ = t.knn(q, kk, md)
            assert ref_ikd.same_knn(io.astype(np.int64), do, ir, dr), (k, kk, md)


@needs_ref
def test_oracle_append_matches_reference_ikdtree_driven_loop():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(GOLD), "..", "tools"))
    from make_golden import append_with_tree
    cfg = api.MeshConfig()
    sensor, scans = synth.make_stream("avia", 3, seed=4, n_points=4000)
    frames = [(s["body_full"].astype(np.float64) @ s["R_true"].T + s["t_true"]).astype(np.float32) for s in scans]
    ref = append_with_tree(frames, cfg.points_minimum_scale, cfg.number_of_pts_append_to_map)
    o = _oracle_with_frames(frames, [s["t_true"] for s in scans], cfg)
    v, _, _ = o.snapshot()
    assert v.shape == ref.shape and np.array_equal(v, ref)
