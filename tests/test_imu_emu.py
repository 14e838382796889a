"""CPU tier: the device bodies of the IMU front-end (immesh_b200/csrc/imu_core.cuh, run by the host emulation harness) against the
oracle restatement of ImuProcess::UndistortPcl (oracle/orc_imu.hpp): propagated state + covariance, IMUpose records and the
time-sorted, motion-compensated cloud -- bit-exact, over several consecutive scans (the members carried from scan to scan included)."""
import numpy as np

import oracle_api as oa
from immesh_b200 import api


def _imu_cfg(rng):
    ang = 0.02
    Rz = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1.0]])
    return dict(cov_gyr=[0.1, 0.12, 0.09], cov_acc=[0.4, 0.5, 0.45], cov_bias_gyr=[1e-4, 1e-4, 2e-4], cov_bias_acc=[1e-3, 2e-3, 1e-3],
                mean_acc_norm=9.78, lid_R=Rz, lid_T=[0.04165, 0.02326, -0.0284])


def _make_scan(rng, k, n, t0, imu_rate=200.0, scan_dt=0.1, dup_stamps=False):
    """scan k: points with time offsets (ms) in [0, 100), IMU samples covering it; hand-held-like motion."""
    beg = t0 + k * scan_dt
    tt = np.arange(beg + 0.5 / imu_rate, beg + scan_dt + 0.5 / imu_rate, 1.0 / imu_rate)
    gyr = 0.3 * np.sin(2.0 * tt)[:, None] * np.array([0.3, -0.2, 1.0]) + rng.normal(0, 0.01, (len(tt), 3))
    acc = np.array([0.2, -0.1, 9.78]) + 0.5 * np.cos(1.5 * tt)[:, None] * np.array([1.0, 0.5, 0.1]) + rng.normal(0, 0.05, (len(tt), 3))
    imu = np.concatenate([tt[:, None], gyr, acc], axis=1)
    curv = rng.uniform(0.0, scan_dt * 1000.0, n).astype(np.float32)
    if dup_stamps:
        curv = np.round(curv / 0.5).astype(np.float32) * np.float32(0.5)       # many equal stamps: the order inside a stamp must be stable
    curv[rng.integers(0, n)] = 0.0                                            # a point at t = 0 stays uncompensated
    xyz = rng.normal(0, 10, (n, 3)).astype(np.float32)
    return beg, imu, np.concatenate([xyz, curv[:, None]], axis=1).astype(np.float32)


def _run(emu_lib, n_scans, n_pts, seed, dup_stamps=False):
    rng = np.random.default_rng(seed)
    cfg = _imu_cfg(rng)
    lio = api.Lio(api.AVIA, lib=emu_lib)
    s0 = lio.get_state()
    s0[12:15] = [0.5, -0.2, 0.05]
    s0[21:24] = [0.0, 0.0, -9.81]
    s0[15:18] = [0.001, -0.002, 0.0005]
    s0[18:21] = [0.01, 0.02, -0.01]
    lio.set_state(s0)
    g, o = api.Imu(cfg, lib=emu_lib), oa.OracleImu(cfg)
    t0 = 100.0
    last = np.array([t0 - 0.002, 0.01, 0.0, 0.02, 0.1, 0.0, 9.8])
    for h in (g, o):
        h.reset(last, t0 - 0.001, 0.0, [0.0, 0.0, 0.0], [0.0, 0.0, 0.0])
    st_o = s0.copy()
    for k in range(n_scans):
        beg, imu, pts = _make_scan(rng, k, n_pts, t0, dup_stamps=dup_stamps)
        out_g = g.undistort(lio, imu, pts, beg)
        st_o, out_o, poses_o = o.undistort(st_o, imu, pts, beg)
        assert np.array_equal(g.poses(), poses_o), f"scan {k}: IMUpose"
        assert np.array_equal(lio.get_state(), st_o), f"scan {k}: propagated state / covariance"
        assert np.array_equal(out_g, out_o), f"scan {k}: compensated cloud"
        assert np.all(np.diff(out_g[:, 3]) >= 0)
        moved = np.abs(out_g[:, :3] - pts[np.argsort(pts[:, 3], kind="stable")][:, :3]).max(axis=1)
        assert np.all(moved[out_g[:, 3] == 0.0] == 0.0)
        if n_pts > 100:
            assert moved.max() > 1e-3
    return lio


def test_undistort_stream_bit_exact(emu_lib):
    _run(emu_lib, 4, 3000, seed=1)


def test_equal_time_stamps_keep_input_order(emu_lib):
    _run(emu_lib, 2, 4000, seed=2, dup_stamps=True)


def test_single_point_and_tiny_scans(emu_lib):
    _run(emu_lib, 3, 1, seed=3)
    _run(emu_lib, 2, 7, seed=4)
