"""GPU tier: the CUDA localization path through the C ABI against the oracle on the same seeded scans."""
import numpy as np
import pytest

from immesh_b200 import api, synth
from lio_common import run_stream_parity, init_velocity
from oracle_api import OracleLio

pytestmark = pytest.mark.gpu


def test_avia_stream_bit_exact(cuda_lib):
    stats = run_stream_parity(cuda_lib, "avia", api.AVIA, n_scans=12, seed=0)
    assert stats[-1]["n_match"] > stats[0]["n_match"]
    assert all(s["pos_err"] < 0.05 for s in stats)


def test_hdl64_stream_bit_exact(cuda_lib):
    stats = run_stream_parity(cuda_lib, "hdl64", api.VELODYNE, n_scans=4, seed=1, n_points=32768)
    assert stats[-1]["n_match"] > 500


def test_fused_step_equals_staged_calls(cuda_lib):
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", 5, seed=3, ext_T=cfg.ext_T)
    a, b = api.Lio(cfg, lib=cuda_lib), api.Lio(cfg, lib=cuda_lib)
    for h in (a, b):
        h.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        init_velocity(h, sensor, scans)
        h.voxel_map_init(scans[0]["body_full"])
    for k in range(1, 5):
        body, dt = scans[k]["body_ds"], scans[k]["dt"]
        a.predict(dt)
        a.lio_state_estimation(body)
        a.map_incremental_grow()
        s, it = b.step(body, dt)
        assert np.array_equal(s, a.get_state())
    assert np.array_equal(a.dump_map(), b.dump_map())


def test_residual_build_dropin(cuda_lib):
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", 4, seed=5, ext_T=cfg.ext_T)
    g, o = api.Lio(cfg, lib=cuda_lib), OracleLio(cfg)
    for h in (g, o):
        h.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        h.voxel_map_init(scans[0]["body_full"])
    for k in (1, 2):
        g.lio_state_estimation(scans[k]["body_ds"]); g.map_incremental_grow()
        o.lio_state_estimation(scans[k]["body_ds"]); o.map_incremental_grow(scans[k]["body_ds"])
    il_g, v_g = g.residual_build(scans[3]["body_ds"])
    il_o, v_o = o.residual_list(scans[3]["body_ds"])
    assert np.array_equal(il_g, il_o)
    assert np.array_equal(v_g, v_o)
    assert len(il_g) > 100


def test_empty_and_tiny_scans(cuda_lib):
    cfg = api.AVIA
    g = api.Lio(cfg, lib=cuda_lib)
    g.voxel_map_init(np.zeros((0, 3), np.float32))
    assert g.counts()["roots"] == 0
    s0 = g.get_state()
    it = g.lio_state_estimation(np.zeros((0, 3), np.float32))
    g.map_incremental_grow()
    pts = np.array([[5.0, 0.1, 0.2], [5.0, 0.2, -0.1], [5.1, 0.0, 0.0]], np.float32)
    g.voxel_map_init(pts)
    assert g.counts()["roots"] >= 1 and g.counts()["err"] == 0
