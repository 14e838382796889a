"""Debug helper: run the avia parity stream on the GPU and print the first rows of the map dump that differ from the oracle."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from immesh_b200 import api, synth
from oracle_api import OracleLio
from lio_common import init_velocity

def main(n_scans=12, seed=0):
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", n_scans, seed=seed, leaf=cfg.filter_size_surf, ext_T=cfg.ext_T)
    lib = api.load_library(os.environ.get('IMMESH_DEBUG_LIB'))
    g, o = api.Lio(cfg, lib=lib), OracleLio(cfg, sum_mode=0)
    for h in (g, o):
        h.set_pose(scans[0]["R_true"], scans[0]["t_true"]); init_velocity(h, sensor, scans); h.voxel_map_init(scans[0]["body_full"])
    for k in range(1, n_scans):
        body = scans[k]["body_ds"]
        g.predict(scans[k]["dt"]); o.predict(scans[k]["dt"])
        ig, io = g.lio_state_estimation(body), o.lio_state_estimation(body)
        se = np.array_equal(g.get_state(), o.get_state())
        g.map_incremental_grow(); o.map_incremental_grow(body)
        dg, do = g.dump_map(), o.dump_map()
        ok = dg.shape == do.shape and np.array_equal(dg, do)
        print(f"scan {k}: iters {ig}/{io} state_equal={se} rows {dg.shape[0]}/{do.shape[0]} map_equal={ok}", flush=True)
        if not ok:
            if dg.shape == do.shape:
                bad = np.where(np.any(dg != do, axis=1))[0]
                print("differing rows:", len(bad))
                for r in bad[:6]:
                    cols = np.where(dg[r] != do[r])[0]
                    print(" row", r, "cols", cols[:12], "gpu", dg[r, :11], "orc", do[r, :11])
            else:
                kg = {tuple(r[:4]) for r in dg}; ko = {tuple(r[:4]) for r in do}
                print("only gpu:", list(kg - ko)[:5], "only oracle:", list(ko - kg)[:5])
            return 1
    print("all equal", g.counts())
    return 0

if __name__ == "__main__":
    sys.exit(main())
