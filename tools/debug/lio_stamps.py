#!/usr/bin/env python
"""Latency budget of one IESKF iteration from the in-kernel clock64 stamps of the debug library (tools/debug/build_stamps.sh).
    python tools/debug/lio_stamps.py C100k 24
Stamps are those of the LAST launch of each kernel (the final iteration of the final scan), thread 0 of block 0; the IESKF update's are
thread 0 of the block that ran it."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import bench  # noqa: E402
from immesh_b200 import api  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C100k"
n_scans = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = api.load_library(os.path.join(ROOT, "tools", "debug", "libimmesh_stamps.so"))
wl = bench.workloads()[name]
scans = bench.get_stream(wl, n_scans)
g = api.Lio(wl["lio"], lib=lib)
g.set_state(bench.init_state_vec(scans))
g.voxel_map_init(scans[0]["body_full"])
for k in range(1, n_scans):
    s, it = g.step(scans[k]["body_ds"], scans[k]["dt"])
st = (C.c_longlong * 64)()
assert lib.immesh_debug_stamps(st) == 0
s = list(st)
d = lambda a, b: s[b] - s[a]
print(f"{name}: n {scans[-1]['body_ds'].shape[0]}, iterations {it}  (cycles; 1000 cycles = 0.52 us at 1.92 GHz)")
print(f"k_match  block 0 thread 0: total {d(0, 7)} | stop+dyn loads {d(0, 1)}  state staged {d(1, 2)}  world point+cov {d(2, 3)}  key+hash+root {d(3, 4)}  "
      f"walk (lane 0) {d(4, 5)}  merge {d(5, 6)}  store+exit {d(6, 7)}")
print(f"k_terms  block 0 thread 0: to staged {d(10, 11)}  terms loop {d(11, 12)}  reduce+atomics {d(12, 13)}  last-block detect {d(13, 14)}")
print(f"IESKF update (thread 0 of its block): total {d(20, 27)} | sums->H^T H, state (-) {d(20, 21)}  B = I + A P11 {d(21, 22)}  LU {d(22, 23)}  "
      f"columns {d(23, 24)}  K1, G6, solution {d(24, 25)}  state (+) {d(25, 26)}  covariance {d(26, 27)}")
