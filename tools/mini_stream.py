#!/usr/bin/env python
"""A few scans through both pipelines with the blocking API: small target for `ncu -k regex:...` captures.
    python tools/mini_stream.py [n_scans] [kind]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from immesh_b200 import api, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
kind = sys.argv[2] if len(sys.argv) > 2 else "avia100k"
cfg = api.AVIA
sensor, scans = synth.make_stream(kind, n + 1, seed=0, leaf=cfg.filter_size_surf, ext_T=cfg.ext_T)
lio, mesh = api.Lio(cfg), api.Mesh(api.MeshConfig())
lio.set_pose(scans[0]["R_true"], scans[0]["t_true"])
lio.voxel_map_init(scans[0]["body_full"])
vg = api.VoxelGrid(1 << 20)
for sc in scans[1:]:
    ds = vg.filter(sc["body_full"], cfg.filter_size_surf)
    s, it = lio.step(ds, sc["dt"])
    mesh.push_frame_from_lio(lio, sc["body_full"])
import time
full = scans[-1]["body_full"]
for on_dev in (False,):
    vg.filter(full, cfg.filter_size_surf, fetch=False)
    t0 = time.perf_counter()
    for _ in range(50):
        m = vg.filter(full, cfg.filter_size_surf, fetch=False)
    dt = (time.perf_counter() - t0) / 50
    print(f"voxelgrid front-end: {len(full)} -> {m} points, {dt * 1e3:.3f} ms per call (host input: memcpy + H2D + 20 kernels + D2H of the count)")
print("ok", lio.counts(), mesh.counts(), "ds points", len(ds))

import ctypes as C
st = (C.c_longlong * 16)()
lib = api.load_library()
if hasattr(lib, "immesh_debug_inverse_stamps"):
    lib.immesh_debug_inverse_stamps.argtypes = [C.c_void_p]
    lib.immesh_debug_inverse_stamps(st)
    s = list(st)
    print("k_pinv cycles: kernel", s[6] - s[5], "| load+enter", s[0] - s[5], "init", s[1] - s[0], "step0", s[2] - s[1], "steps1-17", s[3] - s[2], "back-subst", s[4] - s[3], "store+exit", s[6] - s[4])
    print("step 5 (thread 0): pivot", s[9] - s[8], "barrier1", s[10] - s[9], "read+div(idle for t0)", s[11] - s[10], "barrier2", s[12] - s[11], "update", s[13] - s[12], "barrier3", s[14] - s[13])
