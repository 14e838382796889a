"""Shared driver for the localization parity tests: runs the same scan stream through an implementation
under test (CUDA library or host emulation of the kernels) and the oracle, comparing after every stage."""
import numpy as np

from immesh_b200 import api, synth
from oracle_api import OracleLio

SHAPE_COLS = slice(0, 11)  # key, path, layer, flags, point counters: the integer part of the canonical dump


def init_velocity(h, sensor, scans):
    """Start the constant-velocity filter at the true velocity (as a converged filter would be)."""
    s = h.get_state()
    s[12:15] = (scans[1]["t_true"] - scans[0]["t_true"]) / scans[0]["dt"]
    h.set_state(s)


def run_stream_parity(lib, kind, cfg, n_scans, seed, exact_state=True, n_points=None):
    sensor, scans = synth.make_stream(kind, n_scans, seed=seed, leaf=cfg.filter_size_surf, ext_T=cfg.ext_T, n_points=n_points)
    g = api.Lio(cfg, lib=lib)
    o = OracleLio(cfg, sum_mode=0)
    for h in (g, o):
        h.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        init_velocity(h, sensor, scans)
        h.voxel_map_init(scans[0]["body_full"])
    dg, do = g.dump_map(), o.dump_map()
    assert dg.shape == do.shape
    assert np.array_equal(dg[:, SHAPE_COLS], do[:, SHAPE_COLS]), "octree shape after buildVoxelMap"
    assert np.array_equal(dg, do), "plane parameters after buildVoxelMap"
    stats = []
    for k in range(1, n_scans):
        body = scans[k]["body_ds"]
        g.predict(scans[k]["dt"])
        o.predict(scans[k]["dt"])
        assert np.array_equal(g.get_state(), o.get_state()), f"scan {k}: prediction"
        ig = g.lio_state_estimation(body)
        io = o.lio_state_estimation(body)
        assert ig == io, f"scan {k}: iteration count {ig} vs {io}"
        assert np.array_equal(g.matches(), o.matches()), f"scan {k}: matched point set / layers"
        for it in range(io):
            a, b = g.iter_stats(it), o.iter_stats(it)
            assert a["n_match"] == b["n_match"]
            assert np.array_equal(a["HTH"], b["HTH"]) and np.array_equal(a["HTz"], b["HTz"]), f"scan {k} iter {it}: normal equations"
        sg, so = g.get_state(), o.get_state()
        if exact_state:
            assert np.array_equal(sg, so), f"scan {k}: state max diff {np.abs(sg - so).max()}"
        else:
            assert np.allclose(sg, so, rtol=1e-9, atol=1e-12)
        g.map_incremental_grow()
        o.map_incremental_grow(body)
        dg, do = g.dump_map(), o.dump_map()
        assert dg.shape == do.shape, f"scan {k}: node count {dg.shape[0]} vs {do.shape[0]}"
        assert np.array_equal(dg[:, SHAPE_COLS], do[:, SHAPE_COLS]), f"scan {k}: octree shape"
        assert np.array_equal(dg, do), f"scan {k}: plane parameters"
        stats.append(dict(n=body.shape[0], n_match=g.iter_stats(0)["n_match"], iters=ig,
                          pos_err=float(np.linalg.norm(sg[9:12] - scans[k]["t_true"]))))
    assert g.counts()["err"] == 0
    return stats
