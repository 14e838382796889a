"""GPU tier: the CUDA mesher through the C ABI against the oracle (vertex ids, facet connectivity, flags: bit-exact)."""
import numpy as np
import pytest

from immesh_b200 import api
from mesh_common import run_mesh_parity, world_scans, SMALL
from oracle_api import OracleMesh

pytestmark = pytest.mark.gpu


def test_avia_frames_bit_exact(cuda_lib):
    g, o, stats = run_mesh_parity(cuda_lib, "avia", 8, seed=0)
    assert stats[-1]["n_triangles"] > 10000 and stats[-1]["frame_removed"] > 0


def test_hdl64_frames_bit_exact(cuda_lib):
    run_mesh_parity(cuda_lib, "hdl64", 4, seed=2)


def test_step_one_dense_append(cuda_lib):
    run_mesh_parity(cuda_lib, "avia", 3, seed=4, n_points=9000)


def test_knn_matches_oracle(cuda_lib):
    g, o, _ = run_mesh_parity(cuda_lib, "avia", 3, seed=6)
    v, _, _ = o.snapshot()
    rng = np.random.default_rng(0)
    q = np.concatenate([v[rng.integers(0, len(v), 300)] + rng.normal(0, 0.05, (300, 3)).astype(np.float32),
                        v[:100], rng.uniform(-50, 50, (50, 3)).astype(np.float32)]).astype(np.float32)
    for k, md in ((1, np.inf), (20, np.inf), (20, 1.0), (5, 0.3)):
        ig, dg = g.knn(q, k, md)
        io, do = o.knn(q, k, md)
        assert np.array_equal(ig, io), (k, md)
        assert np.array_equal(dg, do), (k, md)


def _golden(name):
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name))


def test_golden_reference_ikdtree_append_and_knn(cuda_lib):
    """CUDA vertex append + kNN against fixtures produced by the REFERENCE's own ikd-Tree (tools/make_golden.py)."""
    import ref_ikd
    ga, gk = _golden("ikd_append.npz"), _golden("ikd_knn.npz")
    g = api.Mesh(api.MeshConfig(), lib=cuda_lib)
    for k in range(3):
        g.push_frame(ga[f"frame{k}"], ga["pose_t"][k], k)
    v = g.snapshot()[0]
    assert v.shape == ga["verts"].shape and np.array_equal(v, ga["verts"])
    for c, (k, md) in enumerate(gk["cases"]):
        idx, d2 = g.knn(gk["queries"], int(k), float(md))
        assert ref_ikd.same_knn(idx.astype(np.int64), d2, gk[f"idx{c}"].astype(np.int64), gk[f"d2{c}"]), (k, md)


def test_knn_matches_reference_ikdtree_live(cuda_lib):
    """CUDA kNN against the reference ikd-Tree library itself (oracle/_ref travels with the snapshot)."""
    import ref_ikd
    if not ref_ikd.available():
        pytest.skip("oracle/_ref/libref_ikd.so not present")
    g, o, _ = run_mesh_parity(cuda_lib, "avia", 3, seed=9)
    v = g.snapshot()[0]
    t = ref_ikd.RefIkdTree()
    t.add(v)
    rng = np.random.default_rng(1)
    q = (v[rng.integers(0, len(v), 4000)] + rng.normal(0, 0.2, (4000, 3))).astype(np.float32)
    for k, md in ((20, np.inf), (1, 0.1), (20, 1.0)):
        ig, dg = g.knn(q, k, md)
        ir, dr, _ = t.knn(q, k, md)
        assert ref_ikd.same_knn(ig.astype(np.int64), dg, ir, dr), (k, md)


def test_empty_frame_and_restart(cuda_lib):
    cfg = api.MeshConfig(**SMALL)
    g = api.Mesh(cfg, lib=cuda_lib)
    g.push_frame(np.zeros((0, 3), np.float32), np.zeros(3), 0)
    assert g.counts()["n_vertices"] == 0
    pts = np.array([[1.0, 0, 0], [1.1, 0.1, 0], [1.0, 0.15, 0.05]], np.float32)   # one 0.4 m voxel, three xi-cells
    g.push_frame(pts, np.zeros(3), 1)
    c = g.counts()
    assert c["n_vertices"] == 3 and c["n_triangles"] == 1


def test_pipelined_equals_sequential(cuda_lib):
    """LIO(k+1) overlapped with mesh(k) on two streams must give exactly the results of the blocking calls."""
    from immesh_b200 import synth
    from lio_common import init_velocity
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", 7, seed=8, ext_T=cfg.ext_T)
    res = []
    for mode in ("sync", "async"):
        lio, mesh = api.Lio(cfg, lib=cuda_lib), api.Mesh(api.MeshConfig(**SMALL), lib=cuda_lib)
        lio.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        init_velocity(lio, sensor, scans)
        lio.voxel_map_init(scans[0]["body_full"])
        for k in range(1, 7):
            if mode == "sync":
                lio.step(scans[k]["body_ds"], scans[k]["dt"])
                mesh.push_frame_from_lio(lio, scans[k]["body_full"])
            else:
                lio.step_async(scans[k]["body_ds"], dt=scans[k]["dt"])
                mesh.push_frame_from_lio_async(lio, scans[k]["body_full"])
        if mode == "async":
            lio.wait()
            mesh.wait()
        res.append((lio.get_state(), lio.dump_map(), mesh.snapshot()))
    assert np.array_equal(res[0][0], res[1][0])
    assert np.array_equal(res[0][1], res[1][1])
    for a, b in zip(res[0][2], res[1][2]):
        assert np.array_equal(a, b)
    assert len(res[0][2][1]) > 1000
