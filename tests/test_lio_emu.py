"""CPU tier: the __host__ __device__ bodies of the CUDA localization kernels, executed by the host emulation
harness (tests/emu), against the oracle -- bit-exact voxel keys, octree shape, match sets, normal equations, state."""
import numpy as np

from immesh_b200 import api
from lio_common import run_stream_parity


def test_avia_stream_bit_exact(emu_lib):
    stats = run_stream_parity(emu_lib, "avia", api.AVIA, n_scans=12, seed=0)
    assert stats[-1]["n_match"] > stats[0]["n_match"]          # the map densifies
    assert all(s["pos_err"] < 0.05 for s in stats)             # and the filter tracks the trajectory


def test_hdl64_octree_depth_bit_exact(emu_lib):
    # velodyne.yaml: 3 m root voxels, 4 layers, 1000-point nodes -> exercises cut_octo_tree recursion and freezing
    stats = run_stream_parity(emu_lib, "hdl64", api.VELODYNE, n_scans=4, seed=1, n_points=32768)
    assert stats[-1]["n_match"] > 500
