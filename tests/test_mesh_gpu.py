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


def test_empty_frame_and_restart(cuda_lib):
    cfg = api.MeshConfig(**SMALL)
    g = api.Mesh(cfg, lib=cuda_lib)
    g.push_frame(np.zeros((0, 3), np.float32), np.zeros(3), 0)
    assert g.counts()["n_vertices"] == 0
    pts = np.array([[1.0, 0, 0], [1.1, 0.1, 0], [1.0, 0.15, 0.05]], np.float32)   # one 0.4 m voxel, three xi-cells
    g.push_frame(pts, np.zeros(3), 1)
    c = g.counts()
    assert c["n_vertices"] == 3 and c["n_triangles"] == 1
