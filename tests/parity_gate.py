"""Parity gate shared by the GPU tests and bench.py: the same scans through the CUDA path -- the PIPELINED entry points, i.e. the
CUDA-graph-replayed launch sequences the benchmark times -- and through the CPU oracle; filter state, VoxelMap dump, vertex
list, facet set and orientation flags must be identical.  TEST INFRASTRUCTURE (imports the oracle)."""
import numpy as np

from immesh_b200 import api
from oracle_api import OracleLio, OracleMesh


def init_state_vec(scans):
    s = np.zeros(348)
    s[0:9] = scans[0]["R_true"].reshape(9)
    s[9:12] = scans[0]["t_true"]
    s[12:15] = (scans[1]["t_true"] - scans[0]["t_true"]) / scans[0]["dt"]
    for i in range(18):
        s[24 + i * 18 + i] = 1e-7
    return s


def pipeline_parity(lib, cfg, mesh_cfg, scans, n_scans, blocking=False, dev_inputs=None, oracle_threads=(1, 1)):
    """Runs scans[1 .. n_scans] (scans[0] builds the map).  dev_inputs: optional (d_ds, d_full) lists of torch CUDA tensors ->
    the on-device input form the benchmark's `value` region uses.  Returns a report dict; report["ok"] is the verdict."""
    g, o = api.Lio(cfg, lib=lib), OracleLio(cfg, sum_mode=0, omp_threads=oracle_threads[0])
    gm, om = api.Mesh(mesh_cfg, lib=lib), OracleMesh(mesh_cfg, threads=oracle_threads[1])
    s0 = init_state_vec(scans)
    for h in (g, o):
        h.set_state(s0)
        h.voxel_map_init(scans[0]["body_full"])
    iters_o = []
    for k in range(1, n_scans + 1):
        sc = scans[k]
        if blocking:
            g.step(sc["body_ds"], sc["dt"])
            gm.push_frame_from_lio(g, sc["body_full"])
        elif dev_inputs is not None:
            d_ds, d_full = dev_inputs
            g.step_async(d_ds[k].data_ptr(), d_ds[k].shape[0], sc["dt"], on_device=True)
            gm.push_frame_from_lio_async(g, d_full[k].data_ptr(), d_full[k].shape[0], on_device=True)
        else:
            g.step_async(sc["body_ds"], dt=sc["dt"])
            gm.push_frame_from_lio_async(g, sc["body_full"])
        o.predict(sc["dt"])
        iters_o.append(o.lio_state_estimation(sc["body_ds"]))
        o.map_incremental_grow(sc["body_ds"])
        om.push_frame(o.transform_full(sc["body_full"]), o.get_state()[9:12], k)
    if blocking:
        sg, it_g = g.get_state(), None
    else:
        sg, it_g = g.wait()
        gm.wait()
    so = o.get_state()
    dg, do = g.dump_map(), o.dump_map()
    (vg, tg, fg), (vo, to, fo) = gm.snapshot(), om.snapshot()
    rep = {
        "scans": n_scans,
        "state_equal": bool(np.array_equal(sg, so)),
        "state_max_abs_diff": float(np.abs(sg - so).max()),
        "iters_last_equal": (it_g is None) or (it_g == iters_o[-1]),
        "map_rows": [int(dg.shape[0]), int(do.shape[0])],
        "map_equal": bool(dg.shape == do.shape and np.array_equal(dg, do)),
        "vertices": [int(vg.shape[0]), int(vo.shape[0])],
        "vertices_equal": bool(vg.shape == vo.shape and np.array_equal(vg, vo)),
        "facets": [int(tg.shape[0]), int(to.shape[0])],
        "facets_equal": bool(tg.shape == to.shape and np.array_equal(tg, to)),
        "flips_equal": bool(fg.shape == fo.shape and np.array_equal(fg, fo)),
        "lio_err": g.counts()["err"],
    }
    rep["ok"] = all(rep[k] for k in ("state_equal", "iters_last_equal", "map_equal", "vertices_equal", "facets_equal", "flips_equal")) and rep["lio_err"] == 0
    g.close(); gm.close()
    return rep
