"""CPU tier: parameter sweep of the kernel bodies (host emulation) against the oracle -- configurations the shipped yaml files do not
cover: voxel sizes that are not exactly representable (the insert / lookup voxel-size type mismatch of voxel_mapping.cpp:118-127 vs
:172-181 then matters), shallow and deep octrees, small node capacities (early freezing), the dense-Ouster-like 0.2 m setting of
BASELINE config C5, mesh voxel / xi-cell sizes other than 0.4 / 0.1 m, tiny append targets.  Everything bit-exact."""
import dataclasses

import pytest

from immesh_b200 import api
from lio_common import run_stream_parity
from mesh_common import run_mesh_parity

LIO_CASES = [
    dict(voxel_size=0.4, max_layer=2, max_points_size=100, filter_size_surf=0.4),                      # 0.4f != 0.4: border cells differ between insert and lookup
    dict(voxel_size=0.2, max_layer=1, max_points_size=50, filter_size_surf=0.2),                       # C5-like: 0.2 m root voxels, one layer
    dict(voxel_size=1.0, max_layer=3, max_points_size=30, layer_init_size=(5, 5, 5, 5, 5)),            # deep tree, nodes freeze after 30 points
    dict(voxel_size=2.0, max_layer=4, max_points_size=200, dept_err=0.04, beam_err=0.1, max_iteration=3),
    dict(voxel_size=0.5, max_layer=0, max_points_size=100, min_eigen_value=0.005),                     # root voxels only
    dict(voxel_size=0.5, max_layer=2, max_points_size=100, ext_T=(0.0, 0.0, 0.0), max_iteration=2),
]


@pytest.mark.parametrize("case", range(len(LIO_CASES)))
def test_lio_parameter_sweep_bit_exact(emu_lib, case):
    cfg = dataclasses.replace(api.AVIA, **LIO_CASES[case])
    kind = "hdl64" if case % 2 else "avia"
    stats = run_stream_parity(emu_lib, kind, cfg, n_scans=5, seed=20 + case, n_points=16000)
    assert stats[-1]["n_match"] > 50


MESH_CASES = [
    dict(points_minimum_scale=0.05, voxel_resolution=0.2),        # C5-like mesh voxels (xi scaled)
    dict(points_minimum_scale=0.1, voxel_resolution=0.8),         # many vertices per voxel -> block-level triangulation sizes
    dict(points_minimum_scale=0.2, voxel_resolution=0.4),         # coarse xi-grid, sparse vertices, many tiny dilated sets
    dict(points_minimum_scale=0.1, voxel_resolution=0.4, number_of_pts_append_to_map=500),   # large step: few candidates per frame
]


@pytest.mark.parametrize("case", range(len(MESH_CASES)))
def test_mesh_parameter_sweep_bit_exact(emu_lib, case):
    g, o, stats = run_mesh_parity(emu_lib, "hdl64" if case % 2 else "avia", 4, seed=30 + case, cfg_kw=MESH_CASES[case], n_points=16000)
    assert stats[-1]["n_vertices"] > 100
