"""GPU tier: the device pcl::VoxelGrid front-end (immesh_voxelgrid_*) against the oracle restatement
(oracle/orc_frontend.hpp): leaf set, leaf order and float centroids bit-exact; PCL's edge branches."""
import numpy as np
import pytest

import oracle_api as oa
from immesh_b200 import api, synth

pytestmark = pytest.mark.gpu


def _check(vg, pts, leaf):
    got = vg.filter(pts, leaf)
    ref, small, _ = oa.voxel_grid(pts, leaf)
    assert vg.leaf_too_small == small
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.array_equal(got, ref)
    return got


@pytest.mark.parametrize("kind,n_points,leaf", [("avia", None, 0.4), ("hdl64", None, 0.5), ("avia100k", None, 0.4), ("avia", 6000, 0.2)])
def test_scan_shapes_bit_exact(cuda_lib, kind, n_points, leaf):
    vg = api.VoxelGrid(1 << 18, lib=cuda_lib)
    sensor, scans = synth.make_stream(kind, 2, seed=5, n_points=n_points)
    for sc in scans:
        out = _check(vg, sc["body_full"], leaf)
        assert 0 < len(out) < len(sc["body_full"])


def test_edge_cases(cuda_lib):
    vg = api.VoxelGrid(1 << 16, lib=cuda_lib)
    assert vg.filter(np.zeros((0, 3), np.float32), 0.4).shape == (0, 3)
    _check(vg, np.array([[1.5, -2.25, 0.75]], np.float32), 0.4)                      # a single point
    rng = np.random.default_rng(0)
    same = (rng.uniform(0.01, 0.39, (5000, 3)) + np.array([0.4, -0.8, 1.2])).astype(np.float32)
    assert len(_check(vg, same, 0.4)) == 1                                            # one leaf, 5000-term float sum in scan order
    p = rng.normal(0, 3, (30000, 3)).astype(np.float32)
    p[::7] = np.nan
    p[5::11, 1] = np.inf
    _check(vg, p, 0.25)                                                               # non-finite points are skipped
    _check(vg, np.full((10, 3), np.nan, np.float32), 0.4)                             # nothing finite -> empty
    wide = rng.uniform(-3000, 3000, (2000, 3)).astype(np.float32)
    out = _check(vg, wide, 0.01)                                                      # PCL: leaf too small -> output = input
    assert vg.leaf_too_small and np.array_equal(out, wide)
    _check(vg, rng.uniform(-50, 50, (60000, 3)).astype(np.float32), 2.0)              # many points per leaf, ragged last tile
    _check(vg, (rng.integers(-40, 40, (20000, 3)) * 0.5).astype(np.float32), 0.5)     # points exactly on leaf borders / negative cells


def test_device_output_feeds_localization(cuda_lib):
    """filter() on the device, result handed to immesh_lio_step_async as a device pointer: same state as the host-array path."""
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", 4, seed=12, ext_T=cfg.ext_T)
    vg = api.VoxelGrid(1 << 18, lib=cuda_lib)
    states = []
    for mode in ("host", "device"):
        lio = api.Lio(cfg, lib=cuda_lib)
        lio.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        lio.voxel_map_init(scans[0]["body_full"])
        for sc in scans[1:]:
            if mode == "host":
                ds, _, _ = oa.voxel_grid(sc["body_full"], cfg.filter_size_surf)
                lio.step_async(ds, dt=sc["dt"])
            else:
                m = vg.filter(sc["body_full"], cfg.filter_size_surf, fetch=False)
                lio.step_async(vg.device_points(), m, sc["dt"], on_device=True)
            s, _ = lio.wait()
        states.append(s)
    assert np.array_equal(states[0], states[1])


def test_kitti_calib_and_strided_input_bit_exact(cuda_lib):
    """immesh_frontend_prepare: KITTI laser calibration (voxel_mapping.cpp:1844-1859) + repack of [n][3] and [n][4] clouds."""
    sensor, scans = synth.make_stream("hdl64", 1, seed=11, n_points=65536)
    pts = scans[0]["body_full"]
    vg = api.VoxelGrid(1 << 17, lib=cuda_lib)
    ref = oa.kitti_calib(pts)
    assert np.array_equal(vg.prepare(pts, calib_laser=True), ref)
    p4 = np.concatenate([pts, np.arange(len(pts), dtype=np.float32)[:, None]], axis=1)      # xyz + curvature, as the IMU stage emits
    assert np.array_equal(vg.prepare(p4, calib_laser=True), ref)
    assert np.array_equal(vg.prepare(p4, calib_laser=False), pts)
    # chained on the device: calibrated cloud -> VoxelGrid
    ds = vg.filter(vg.input_points(), 0.5, n=len(pts), on_device=True)
    vg.prepare(pts, calib_laser=True, fetch=False)
    ds = vg.filter(vg.input_points(), 0.5, n=len(pts), on_device=True)
    ref_ds, _, _ = oa.voxel_grid(ref, 0.5)
    assert np.array_equal(ds, ref_ds)


def test_device_resident_frontend_chain_equals_staged_calls(cuda_lib):
    """immesh_lio_step_async_raw: raw scan -> [calibration] -> VoxelGrid -> localization without a host round trip of the down-sampled
    cloud or its size, against the same steps through the host (oracle front-end + oracle localization)."""
    import dataclasses
    from oracle_api import OracleLio
    from lio_common import init_velocity
    cfg = dataclasses.replace(api.VELODYNE, calib_laser=1)
    sensor, scans = synth.make_stream("hdl64", 4, seed=12, leaf=cfg.filter_size_surf, n_points=65536)
    g, o = api.Lio(cfg, lib=cuda_lib), OracleLio(cfg)
    vg = api.VoxelGrid(1 << 17, lib=cuda_lib)
    for h in (g, o):
        h.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        init_velocity(h, sensor, scans)
        h.voxel_map_init(oa.kitti_calib(scans[0]["body_full"]))
    for k in (1, 2, 3):
        raw = scans[k]["body_full"]
        vg.step_async_raw(g, raw, cfg.filter_size_surf, dt=scans[k]["dt"], calib_laser=True)
        ds, _, _ = oa.voxel_grid(oa.kitti_calib(raw), cfg.filter_size_surf)
        o.predict(scans[k]["dt"])
        o.lio_state_estimation(ds)
        o.map_incremental_grow(ds)
    s, it = g.wait()
    assert np.array_equal(s, o.get_state())
    assert np.array_equal(g.dump_map(), o.dump_map())
