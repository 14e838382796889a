"""CPU tier: the __host__ __device__ bodies of the CUDA mesher (host emulation) against the oracle."""
import numpy as np

from mesh_common import run_mesh_parity


def test_avia_frames_bit_exact(emu_lib):
    g, o, stats = run_mesh_parity(emu_lib, "avia", 6, seed=0)
    assert stats[-1]["n_triangles"] > 10000 and stats[-1]["frame_removed"] > 0   # re-meshing happens


def test_hdl64_frames_bit_exact(emu_lib):
    g, o, stats = run_mesh_parity(emu_lib, "hdl64", 3, seed=2, n_points=65536)
    assert stats[-1]["n_vertices"] > 5000


def test_step_one_dense_append(emu_lib):
    # fewer points than the append target: step = 1, long same-surface conflict chains in the greedy vertex insertion
    g, o, stats = run_mesh_parity(emu_lib, "avia", 3, seed=4, n_points=9000)
    assert stats[0]["frame_new_vertices"] > 1000
