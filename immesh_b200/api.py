"""ctypes binding of include/immesh_b200.h (test / bench harness; no compute happens in Python)."""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libimmesh_b200.so")
_lib_cache: dict[str, C.CDLL] = {}

STATE_DOUBLES = 348


class _LioCfg(C.Structure):
    _fields_ = [
        ("voxel_size", C.c_double), ("max_layer", C.c_int), ("layer_init_size", C.c_int * 5), ("max_points_size", C.c_int),
        ("min_eigen_value", C.c_double), ("dept_err", C.c_double), ("beam_err", C.c_double),
        ("ext_R", C.c_double * 9), ("ext_T", C.c_double * 3), ("max_iteration", C.c_int), ("calib_laser", C.c_int),
        ("hash_capacity_log2", C.c_int), ("max_nodes", C.c_int), ("max_chunks", C.c_int), ("max_scan_points", C.c_int),
    ]


class _MeshCfg(C.Structure):
    _fields_ = [
        ("points_minimum_scale", C.c_double), ("voxel_resolution", C.c_double), ("number_of_pts_append_to_map", C.c_int),
        ("max_vertices", C.c_int), ("max_triangles", C.c_int), ("max_voxels", C.c_int), ("max_frame_points", C.c_int),
    ]


class _ImuCfg(C.Structure):
    _fields_ = [("cov_gyr", C.c_double * 3), ("cov_acc", C.c_double * 3), ("cov_bias_gyr", C.c_double * 3), ("cov_bias_acc", C.c_double * 3),
                ("mean_acc_norm", C.c_double), ("lid_R", C.c_double * 9), ("lid_T", C.c_double * 3), ("max_points", C.c_int), ("max_imu", C.c_int)]


@dataclasses.dataclass
class LioConfig:
    """Hot-path parameters of Voxel_mapping (config/*.yaml of the reference)."""
    voxel_size: float = 0.5
    max_layer: int = 2
    layer_init_size: tuple = (5, 5, 5, 5, 5)
    max_points_size: int = 100
    min_eigen_value: float = 0.01
    dept_err: float = 0.02
    beam_err: float = 0.05
    ext_R: tuple = (1, 0, 0, 0, 1, 0, 0, 0, 1)
    ext_T: tuple = (0.04165, 0.02326, -0.0284)
    max_iteration: int = 4
    calib_laser: int = 0
    filter_size_surf: float = 0.4   # leaf of the down-sampling step before the path
    hash_capacity_log2: int = 0
    max_nodes: int = 0
    max_chunks: int = 0
    max_scan_points: int = 0


@dataclasses.dataclass
class MeshConfig:
    points_minimum_scale: float = 0.1
    voxel_resolution: float = 0.4
    number_of_pts_append_to_map: int = 10000
    max_vertices: int = 0
    max_triangles: int = 0
    max_voxels: int = 0
    max_frame_points: int = 0


AVIA = LioConfig()  # config/avia.yaml
VELODYNE = LioConfig(voxel_size=3.0, max_layer=4, max_points_size=1000, dept_err=0.04, beam_err=0.1, ext_T=(0, 0, 0),
                     max_iteration=3, calib_laser=0, filter_size_surf=0.5)  # config/velodyne.yaml (calib_laser applied upstream)


def load_library(path: Optional[str] = None) -> C.CDLL:
    """Load the C-ABI shared library.  There is no fallback: a missing CUDA build is an error."""
    path = path or _LIB_PATH
    if path in _lib_cache:
        return _lib_cache[path]
    if not os.path.exists(path):
        raise RuntimeError(f"{path} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()')")
    lib = C.CDLL(path)
    vp, ip, dp, fp = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)
    lib.immesh_lio_create.argtypes = [C.POINTER(_LioCfg), C.POINTER(vp)]
    lib.immesh_lio_destroy.argtypes = [vp]
    lib.immesh_lio_set_state.argtypes = [vp, dp]
    lib.immesh_lio_get_state.argtypes = [vp, dp]
    lib.immesh_lio_predict.argtypes = [vp, C.c_double, C.c_double, C.c_double]
    lib.immesh_voxelmap_build.argtypes = [vp, fp, C.c_int]
    lib.immesh_lio_estimate.argtypes = [vp, fp, C.c_int, ip]
    lib.immesh_voxelmap_update.argtypes = [vp]
    lib.immesh_lio_iter_stats.argtypes = [vp, C.c_int, dp]
    lib.immesh_lio_matches.argtypes = [vp, ip, C.c_int]
    lib.immesh_voxelmap_dump.argtypes = [vp, dp, C.c_int64]
    lib.immesh_voxelmap_dump.restype = C.c_int64
    lib.immesh_voxelmap_counts.argtypes = [vp, C.POINTER(C.c_int64)]
    for name, args in (
        ("immesh_lio_step", [vp, fp, C.c_int, C.c_double, C.c_double, C.c_double, dp, ip]),
        ("immesh_lio_step_dev", [vp, vp, C.c_int, C.c_double, C.c_double, C.c_double, dp, ip]),
        ("immesh_residual_build", [vp, fp, C.c_int, ip, dp, C.c_int, ip]),
        ("immesh_mesh_push_frame_dev", [vp, vp, C.c_int, dp, C.c_int]),
        ("immesh_mesh_push_frame_from_lio", [vp, vp, vp, C.c_int, C.c_int]),
        ("immesh_lio_match_nodes", [vp, ip, C.c_int]),
        ("immesh_voxelmap_build_pv", [vp, dp, dp, C.c_int]),
        ("immesh_voxelmap_update_pv", [vp, dp, dp, C.c_int]),
        ("immesh_residual_build_pv", [vp, dp, dp, dp, C.c_int, ip, dp, C.c_int, ip]),
        ("immesh_comm_unique_id", [C.c_char_p]),
        ("immesh_lio_shard", [vp, C.c_int, C.c_int, C.c_char_p]),
        ("immesh_mesh_shard", [vp, C.c_int, C.c_int, C.c_char_p]),
        ("immesh_lio_shard_transport", [vp]),
        ("immesh_mesh_shard_transport", [vp]),
        ("immesh_lio_step_async", [vp, vp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double]),
        ("immesh_lio_wait", [vp, dp, ip]),
        ("immesh_lio_enqueue_memset", [vp, vp, C.c_size_t]),
        ("immesh_mesh_push_frame_from_lio_async", [vp, vp, vp, C.c_int, C.c_int]),
        ("immesh_mesh_wait", [vp]),
        ("immesh_pipeline_mark_begin", [vp]),
        ("immesh_pipeline_mark_end", [vp, vp, dp]),
        ("immesh_mesh_work_stats", [vp, C.POINTER(C.c_int64)]),
        ("immesh_host_wait_ms", [vp, vp, dp]),
        ("immesh_lio_work_stats", [vp, C.POINTER(C.c_int64)]),
        ("immesh_profile_enable", [C.c_int]),
        ("immesh_profile_reset", []),
        ("immesh_profile_report", [C.c_char_p, C.c_int]),
        ("immesh_profile_timeline", [C.c_char_p, C.c_int]),
        ("immesh_launch_count", []),
        ("immesh_graph_stats", [vp, vp, C.POINTER(C.c_int64)]),
        ("immesh_lio_last_timing", [vp, dp]),
        ("immesh_mesh_create", [C.POINTER(_MeshCfg), C.POINTER(vp)]),
        ("immesh_mesh_destroy", [vp]),
        ("immesh_mesh_push_frame", [vp, fp, C.c_int, dp, C.c_int]),
        ("immesh_mesh_counts", [vp, C.POINTER(C.c_int64)]),
        ("immesh_mesh_snapshot", [vp, fp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
        ("immesh_knn", [vp, fp, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int32), fp]),
        ("immesh_mesh_last_timing", [vp, dp]),
        ("immesh_write_ply", [C.c_char_p, fp, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]),
        ("immesh_write_pcd", [C.c_char_p, fp, C.c_int]),
        ("immesh_mesh_render_depth", [vp, dp, C.c_int, C.c_int, C.c_double, C.c_double, dp, dp, fp, fp, C.POINTER(C.c_int32), ip]),
        ("immesh_mesh_reconstruct_from_pointcloud", [vp, vp, vp, C.c_int, C.c_int, C.c_double, ip]),
        ("immesh_mesh_smooth_all", [vp, C.c_double, C.c_int, dp]),
        ("immesh_mesh_region_stream", [vp, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), ip]),
        ("immesh_kitti_pose_line", [dp, C.c_double, C.c_char_p, C.c_int]),
        ("immesh_voxelgrid_create", [C.c_int, C.POINTER(vp)]),
        ("immesh_voxelgrid_destroy", [vp]),
        ("immesh_voxelgrid_filter", [vp, vp, C.c_int, C.c_int, C.c_float, vp, ip, ip]),
        ("immesh_voxelgrid_device_points", [vp]),
        ("immesh_frontend_prepare", [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
        ("immesh_voxelgrid_input_points", [vp]),
        ("immesh_lio_step_async_raw", [vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_double, C.c_double, C.c_double]),
        ("immesh_lio_step_async_dev_n", [vp, vp, C.c_int, vp, C.c_double, C.c_double, C.c_double]),
        ("immesh_imu_create", [C.POINTER(_ImuCfg), C.POINTER(vp)]),
        ("immesh_imu_destroy", [vp]),
        ("immesh_imu_reset", [vp, dp, C.c_double, C.c_double, dp, dp]),
        ("immesh_imu_undistort", [vp, vp, dp, C.c_int, vp, C.c_int, C.c_int, C.c_double, vp]),
        ("immesh_imu_device_points", [vp]),
        ("immesh_imu_get_poses", [vp, dp, C.c_int]),
    ):
        if hasattr(lib, name):
            getattr(lib, name).argtypes = args
    if hasattr(lib, "immesh_imu_device_points"):
        lib.immesh_imu_device_points.restype = C.c_void_p
    if hasattr(lib, "immesh_voxelgrid_device_points"):
        lib.immesh_voxelgrid_device_points.restype = C.c_void_p
    if hasattr(lib, "immesh_voxelgrid_input_points"):
        lib.immesh_voxelgrid_input_points.restype = C.c_void_p
    if hasattr(lib, "immesh_launch_count"):
        lib.immesh_launch_count.restype = C.c_longlong
    if hasattr(lib, "immesh_last_error"):
        lib.immesh_last_error.restype = C.c_char_p
    if hasattr(lib, "immesh_version"):
        lib.immesh_version.restype = C.c_char_p
    _lib_cache[path] = lib
    return lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


def _check(lib, rc, what):
    if rc != 0:
        msg = lib.immesh_last_error().decode() if hasattr(lib, "immesh_last_error") else ""
        raise RuntimeError(f"{what} failed with code {rc}: {msg}")


class Lio:
    """Mirror of the reference's Voxel_mapping hot-path methods (voxel_map_init / lio_state_estimation /
    map_incremental_grow) on the device-resident VoxelMap."""

    def __init__(self, cfg: LioConfig = AVIA, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        self.cfg = cfg
        c = _LioCfg()
        c.voxel_size = cfg.voxel_size
        c.max_layer = cfg.max_layer
        for i in range(5):
            c.layer_init_size[i] = cfg.layer_init_size[i]
        c.max_points_size = cfg.max_points_size
        c.min_eigen_value = cfg.min_eigen_value
        c.dept_err, c.beam_err = cfg.dept_err, cfg.beam_err
        for i in range(9):
            c.ext_R[i] = cfg.ext_R[i]
        for i in range(3):
            c.ext_T[i] = cfg.ext_T[i]
        c.max_iteration, c.calib_laser = cfg.max_iteration, cfg.calib_laser
        c.hash_capacity_log2, c.max_nodes, c.max_chunks, c.max_scan_points = cfg.hash_capacity_log2, cfg.max_nodes, cfg.max_chunks, cfg.max_scan_points
        self._h = C.c_void_p()
        _check(self.lib, self.lib.immesh_lio_create(C.byref(c), C.byref(self._h)), "immesh_lio_create")

    def close(self):
        if self._h:
            self.lib.immesh_lio_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- state
    def get_state(self) -> np.ndarray:
        s = np.zeros(STATE_DOUBLES)
        _check(self.lib, self.lib.immesh_lio_get_state(self._h, s.ctypes.data_as(C.POINTER(C.c_double))), "get_state")
        return s

    def set_state(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.size == STATE_DOUBLES
        _check(self.lib, self.lib.immesh_lio_set_state(self._h, s.ctypes.data_as(C.POINTER(C.c_double))), "set_state")

    def set_pose(self, R, t):
        s = self.get_state()
        s[0:9] = np.asarray(R, dtype=np.float64).reshape(9)
        s[9:12] = t
        self.set_state(s)

    # --- reference-named operations
    def predict(self, dt, cov_gyr=0.1, cov_acc=0.1):
        _check(self.lib, self.lib.immesh_lio_predict(self._h, dt, cov_gyr, cov_acc), "predict")

    def voxel_map_init(self, body_full):
        a, p = _f32(body_full)
        _check(self.lib, self.lib.immesh_voxelmap_build(self._h, p, a.shape[0]), "voxelmap_build")

    def lio_state_estimation(self, body_ds) -> int:
        a, p = _f32(body_ds)
        it = C.c_int(0)
        _check(self.lib, self.lib.immesh_lio_estimate(self._h, p, a.shape[0], C.byref(it)), "lio_estimate")
        self._last_n = a.shape[0]
        return it.value

    def map_incremental_grow(self):
        _check(self.lib, self.lib.immesh_voxelmap_update(self._h), "voxelmap_update")

    def step(self, body_ds, dt=0.0, cov_gyr=0.1, cov_acc=0.1):
        a, p = _f32(body_ds)
        it = C.c_int(0)
        s = np.zeros(STATE_DOUBLES)
        _check(self.lib, self.lib.immesh_lio_step(self._h, p, a.shape[0], dt, cov_gyr, cov_acc, s.ctypes.data_as(C.POINTER(C.c_double)), C.byref(it)), "lio_step")
        self._last_n = a.shape[0]
        return s, it.value

    def step_dev(self, dev_ptr, n, dt=0.0, cov_gyr=0.1, cov_acc=0.1):
        """Same as step() with the scan already in device memory (dev_ptr: integer CUDA address of float32[n][3])."""
        it = C.c_int(0)
        s = np.zeros(STATE_DOUBLES)
        _check(self.lib, self.lib.immesh_lio_step_dev(self._h, C.c_void_p(dev_ptr), n, dt, cov_gyr, cov_acc, s.ctypes.data_as(C.POINTER(C.c_double)), C.byref(it)), "lio_step_dev")
        self._last_n = n
        return s, it.value

    def step_async(self, body, n=None, dt=0.0, cov_gyr=0.1, cov_acc=0.1, on_device=False):
        """Queue one scan (predict + estimate + map update) without waiting; body: float32[n][3] host array or CUDA address."""
        if on_device:
            ptr = C.c_void_p(body)
        else:
            a, p = _f32(body)
            n = a.shape[0]
            ptr = C.cast(p, C.c_void_p)
        _check(self.lib, self.lib.immesh_lio_step_async(self._h, ptr, n, 1 if on_device else 0, dt, cov_gyr, cov_acc), "lio_step_async")
        self._last_n = n

    def wait(self):
        it = C.c_int(0)
        s = np.zeros(STATE_DOUBLES)
        _check(self.lib, self.lib.immesh_lio_wait(self._h, s.ctypes.data_as(C.POINTER(C.c_double)), C.byref(it)), "lio_wait")
        return s, it.value

    def enqueue_memset(self, dev_ptr, nbytes):
        _check(self.lib, self.lib.immesh_lio_enqueue_memset(self._h, C.c_void_p(dev_ptr), nbytes), "enqueue_memset")

    def shard(self, rank: int, nranks: int, unique_id: bytes):
        """Shard the VoxelMap over nranks processes (NCCL); unique_id: 128 bytes from comm_unique_id() of rank 0."""
        _check(self.lib, self.lib.immesh_lio_shard(self._h, rank, nranks, unique_id), "lio_shard")

    def shard_transport(self) -> str:
        return ("none", "nccl", "peer-window")[self.lib.immesh_lio_shard_transport(self._h)]

    def residual_build(self, body_ds):
        a, p = _f32(body_ds)
        n = a.shape[0]
        il = np.zeros((n, 2), dtype=np.int32)
        vals = np.zeros((n, 31))
        m = C.c_int(0)
        _check(self.lib, self.lib.immesh_residual_build(self._h, p, n, il.ctypes.data_as(C.POINTER(C.c_int)), vals.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(m)), "residual_build")
        return il[: m.value], vals[: m.value]

    # --- the three free functions of src/voxel_mapping.hpp:80-105 on caller-built Point_with_var lists
    @staticmethod
    def _f64(a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return a, a.ctypes.data_as(C.POINTER(C.c_double))

    def voxelmap_build_pv(self, pts_world, var9):
        (pw, ppw), (v9, pv9) = self._f64(pts_world), self._f64(var9)
        _check(self.lib, self.lib.immesh_voxelmap_build_pv(self._h, ppw, pv9, pw.shape[0]), "voxelmap_build_pv")

    def voxelmap_update_pv(self, pts_world, var9):
        (pw, ppw), (v9, pv9) = self._f64(pts_world), self._f64(var9)
        _check(self.lib, self.lib.immesh_voxelmap_update_pv(self._h, ppw, pv9, pw.shape[0]), "voxelmap_update_pv")

    def residual_build_pv(self, pts_body, pts_world, var9):
        (pb, ppb), (pw, ppw), (v9, pv9) = self._f64(pts_body), self._f64(pts_world), self._f64(var9)
        n = pb.shape[0]
        il = np.zeros((max(n, 1), 2), dtype=np.int32)
        vals = np.zeros((max(n, 1), 31))
        m = C.c_int(0)
        _check(self.lib, self.lib.immesh_residual_build_pv(self._h, ppb, ppw, pv9, n, il.ctypes.data_as(C.POINTER(C.c_int)), vals.ctypes.data_as(C.POINTER(C.c_double)), n, C.byref(m)), "residual_build_pv")
        return il[: m.value], vals[: m.value]

    # --- diagnostics
    def iter_stats(self, it):
        o = np.zeros(63)
        _check(self.lib, self.lib.immesh_lio_iter_stats(self._h, it, o.ctypes.data_as(C.POINTER(C.c_double))), "iter_stats")
        return dict(HTH=o[:36].reshape(6, 6).copy(), HTz=o[36:42].copy(), n_match=int(o[42]), total_residual=o[43], solution=o[44:62].copy(), converged=int(o[62]))

    def matches(self, n=None):
        n = n or self._last_n
        o = np.zeros(n, dtype=np.int32)
        _check(self.lib, self.lib.immesh_lio_matches(self._h, o.ctypes.data_as(C.POINTER(C.c_int)), n), "matches")
        return o

    def match_nodes(self, n=None):
        n = n or self._last_n
        o = np.zeros(n, dtype=np.int32)
        _check(self.lib, self.lib.immesh_lio_match_nodes(self._h, o.ctypes.data_as(C.POINTER(C.c_int)), n), "match_nodes")
        return o

    def dump_map(self) -> np.ndarray:
        rows = self.lib.immesh_voxelmap_dump(self._h, None, 0)
        out = np.zeros((rows, 45))
        r2 = self.lib.immesh_voxelmap_dump(self._h, out.ctypes.data_as(C.POINTER(C.c_double)), rows)
        assert r2 == rows
        return out

    def counts(self):
        o = np.zeros(4, dtype=np.int64)
        self.lib.immesh_voxelmap_counts(self._h, o.ctypes.data_as(C.POINTER(C.c_int64)))
        return dict(roots=int(o[0]), nodes=int(o[1]), chunks=int(o[2]), err=int(o[3]))

    def work_stats(self):
        o = np.zeros(4, dtype=np.int64)
        _check(self.lib, self.lib.immesh_lio_work_stats(self._h, o.ctypes.data_as(C.POINTER(C.c_int64))), "lio_work_stats")
        return dict(n=int(o[0]), roots=int(o[1]), refits=int(o[2]), refit_points=int(o[3]))

    def last_timing(self):
        o = np.zeros(3)
        self.lib.immesh_lio_last_timing(self._h, o.ctypes.data_as(C.POINTER(C.c_double)))
        return o


class Mesh:
    """Mirror of incremental_mesh_reconstruction + Global_map/Triangle_manager snapshot + KD_TREE::Nearest_Search."""

    def __init__(self, cfg: MeshConfig = MeshConfig(), lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        self.cfg = cfg
        c = _MeshCfg(cfg.points_minimum_scale, cfg.voxel_resolution, cfg.number_of_pts_append_to_map, cfg.max_vertices, cfg.max_triangles,
                     cfg.max_voxels, cfg.max_frame_points)
        self._h = C.c_void_p()
        _check(self.lib, self.lib.immesh_mesh_create(C.byref(c), C.byref(self._h)), "immesh_mesh_create")

    def close(self):
        if self._h:
            self.lib.immesh_mesh_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def push_frame(self, world_xyz, pose_t, frame_idx=0):
        a, p = _f32(world_xyz)
        t = np.ascontiguousarray(pose_t, dtype=np.float64)
        _check(self.lib, self.lib.immesh_mesh_push_frame(self._h, p, a.shape[0], t.ctypes.data_as(C.POINTER(C.c_double)), frame_idx), "mesh_push_frame")

    def push_frame_dev(self, dev_ptr, n, pose_t, frame_idx=0):
        t = np.ascontiguousarray(pose_t, dtype=np.float64)
        _check(self.lib, self.lib.immesh_mesh_push_frame_dev(self._h, C.c_void_p(dev_ptr), n, t.ctypes.data_as(C.POINTER(C.c_double)), frame_idx), "mesh_push_frame_dev")

    def push_frame_from_lio(self, lio: "Lio", body_full, n=None, on_device=False):
        """map_incremental_grow's hand-off: transform the full-resolution body scan with lio's converged state on the
        device and mesh it.  body_full: float32[n][3] host array, or an integer CUDA address when on_device."""
        if on_device:
            ptr = C.c_void_p(body_full)
        else:
            a, p = _f32(body_full)
            n = a.shape[0]
            ptr = C.cast(p, C.c_void_p)
        _check(self.lib, self.lib.immesh_mesh_push_frame_from_lio(self._h, lio._h, ptr, n, 1 if on_device else 0), "mesh_push_frame_from_lio")

    def push_frame_from_lio_async(self, lio: "Lio", body_full, n=None, on_device=False):
        if on_device:
            ptr = C.c_void_p(body_full)
        else:
            a, p = _f32(body_full)
            n = a.shape[0]
            ptr = C.cast(p, C.c_void_p)
        _check(self.lib, self.lib.immesh_mesh_push_frame_from_lio_async(self._h, lio._h, ptr, n, 1 if on_device else 0), "mesh_push_frame_from_lio_async")

    def wait(self):
        _check(self.lib, self.lib.immesh_mesh_wait(self._h), "mesh_wait")

    def counts(self):
        o = np.zeros(8, dtype=np.int64)
        _check(self.lib, self.lib.immesh_mesh_counts(self._h, o.ctypes.data_as(C.POINTER(C.c_int64))), "mesh_counts")
        keys = ["n_vertices", "n_triangles", "frame_new_vertices", "frame_voxels_meshed", "frame_added", "frame_removed", "n_voxels", "n_activated"]
        return dict(zip(keys, (int(v) for v in o)))

    def shard(self, rank: int, nranks: int, unique_id: bytes):
        """Shard the per-voxel meshing stage over nranks processes (NCCL, own communicator: pass a second unique id)."""
        _check(self.lib, self.lib.immesh_mesh_shard(self._h, rank, nranks, unique_id), "mesh_shard")

    def shard_transport(self) -> str:
        return ("none", "nccl", "peer-window")[self.lib.immesh_mesh_shard_transport(self._h)]

    def work_stats(self):
        o = np.zeros(8, dtype=np.int64)
        _check(self.lib, self.lib.immesh_mesh_work_stats(self._h, o.ctypes.data_as(C.POINTER(C.c_int64))), "mesh_work_stats")
        keys = ["candidates", "gathered", "queries", "dilated", "faces", "voxels_meshed", "add_entries", "remove_entries"]
        return dict(zip(keys, (int(v) for v in o)))

    def snapshot(self):
        c = self.counts()
        v = np.zeros((c["n_vertices"], 3), dtype=np.float32)
        t = np.zeros((c["n_triangles"], 3), dtype=np.int32)
        f = np.zeros(c["n_triangles"], dtype=np.int32)
        _check(self.lib, self.lib.immesh_mesh_snapshot(self._h, v.ctypes.data_as(C.POINTER(C.c_float)), t.ctypes.data_as(C.POINTER(C.c_int32)),
                                                     f.ctypes.data_as(C.POINTER(C.c_int32))), "mesh_snapshot")
        return v, t, f

    def knn(self, q, k, max_dist=float("inf")):
        a, p = _f32(q)
        nq = a.shape[0]
        idx = np.zeros((nq, k), dtype=np.int32)
        d2 = np.zeros((nq, k), dtype=np.float32)
        _check(self.lib, self.lib.immesh_knn(self._h, p, nq, k, max_dist, idx.ctypes.data_as(C.POINTER(C.c_int32)), d2.ctypes.data_as(C.POINTER(C.c_float))), "knn")
        return idx, d2

    def reconstruct_from_pointcloud(self, vg: "VoxelGrid", pts, minimum_pts_distance: float) -> int:
        """reconstruct_mesh_from_pointcloud: VoxelGrid(leaf = minimum_pts_distance) + one frame with the identity pose; returns the down-sampled size."""
        a, p = _f32(pts)
        m = C.c_int(0)
        _check(self.lib, self.lib.immesh_mesh_reconstruct_from_pointcloud(self._h, vg._h, C.cast(p, C.c_void_p), a.shape[0], 0, minimum_pts_distance, C.byref(m)), "mesh_reconstruct_from_pointcloud")
        return m.value

    def render_depth(self, intrinsics, width, height, z_near, z_far, cam_R, cam_t):
        """(depth float32[h,w] (-1 = empty), points float32[n,3], pixel int32[n]) -- depth rasterisation of the live mesh."""
        K = np.ascontiguousarray(intrinsics, dtype=np.float64)
        R = np.ascontiguousarray(cam_R, dtype=np.float64).reshape(9)
        t = np.ascontiguousarray(cam_t, dtype=np.float64)
        depth = np.zeros((height, width), dtype=np.float32)
        pts = np.zeros((height * width, 3), dtype=np.float32)
        pix = np.zeros(height * width, dtype=np.int32)
        n = C.c_int(0)
        dp = C.POINTER(C.c_double)
        _check(self.lib, self.lib.immesh_mesh_render_depth(self._h, K.ctypes.data_as(dp), width, height, z_near, z_far, R.ctypes.data_as(dp), t.ctypes.data_as(dp),
                                                           depth.ctypes.data_as(C.POINTER(C.c_float)), pts.ctypes.data_as(C.POINTER(C.c_float)),
                                                           pix.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(n)), "mesh_render_depth")
        return depth, pts[:n.value].copy(), pix[:n.value].copy()

    def smooth_all(self, smooth_factor=0.1, knn=20):
        """smooth_all_pts: returns the smoothed positions float64[nv,3] (and stores them as the vertices' smoothed positions)."""
        nv = self.counts()["n_vertices"]
        out = np.zeros((nv, 3))
        _check(self.lib, self.lib.immesh_mesh_smooth_all(self._h, smooth_factor, knn, out.ctypes.data_as(C.POINTER(C.c_double))), "mesh_smooth_all")
        return out

    def region_stream(self, region_size=10.0, cap=1 << 16):
        """(region_keys int32[nr,3], region_offsets int32[nr+1], triangles int32[nt,3]) -- the viewer's region-bucketed triangle sets."""
        nt = self.counts()["n_triangles"]
        keys, offs = np.zeros((cap, 3), dtype=np.int32), np.zeros(cap + 1, dtype=np.int32)
        tri = np.zeros((max(nt, 1), 3), dtype=np.int32)
        nr = C.c_int(0)
        i32 = C.POINTER(C.c_int32)
        _check(self.lib, self.lib.immesh_mesh_region_stream(self._h, region_size, keys.ctypes.data_as(i32), offs.ctypes.data_as(i32), cap, tri.ctypes.data_as(i32), C.byref(nr)), "mesh_region_stream")
        return keys[:nr.value].copy(), offs[:nr.value + 1].copy(), tri[:nt].copy()

    def last_timing(self):
        o = np.zeros(4)
        self.lib.immesh_mesh_last_timing(self._h, o.ctypes.data_as(C.POINTER(C.c_double)))
        return o


class VoxelGrid:
    """pcl::VoxelGrid front-end on the device (immesh_voxelgrid_*): filter(points, leaf) -> centroids, leaves in PCL index order."""

    def __init__(self, max_points: int = 1 << 20, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        self._h = C.c_void_p()
        _check(self.lib, self.lib.immesh_voxelgrid_create(max_points, C.byref(self._h)), "voxelgrid_create")
        self.leaf_too_small = False
        self.m = 0

    def close(self):
        if self._h:
            self.lib.immesh_voxelgrid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def filter(self, pts, leaf: float, n=None, on_device=False, fetch=True):
        """pts: float32[n,3] host array, or a device pointer (int) with n and on_device=True.  Returns float32[m,3] (or m if not fetch)."""
        if on_device:
            ptr, cnt = C.c_void_p(int(pts)), int(n)
        else:
            a = np.ascontiguousarray(pts, dtype=np.float32).reshape(-1, 3)
            ptr, cnt = a.ctypes.data_as(C.c_void_p), a.shape[0]
        out = np.zeros((max(cnt, 1), 3), dtype=np.float32) if fetch else None
        m, small = C.c_int(0), C.c_int(0)
        _check(self.lib, self.lib.immesh_voxelgrid_filter(self._h, ptr, cnt, 1 if on_device else 0, C.c_float(leaf),
                                                          out.ctypes.data_as(C.c_void_p) if fetch else None, C.byref(m), C.byref(small)), "voxelgrid_filter")
        self.m, self.leaf_too_small = m.value, bool(small.value)
        return out[:m.value].copy() if fetch else m.value

    def device_points(self) -> int:
        return int(self.lib.immesh_voxelgrid_device_points(self._h) or 0)

    def prepare(self, pts, calib_laser=False, fetch=True):
        """KITTI laser calibration (optional) + repack of a float32[n,3] or [n,4] cloud; the packed cloud stays on the device
        (input_points()).  Returns it as float32[n,3] when fetch."""
        a = np.ascontiguousarray(pts, dtype=np.float32)
        n, stride = a.shape[0], a.shape[1]
        out = np.zeros((n, 3), dtype=np.float32) if fetch else None
        _check(self.lib, self.lib.immesh_frontend_prepare(self._h, a.ctypes.data_as(C.c_void_p), n, stride, 0, 1 if calib_laser else 0,
                                                          out.ctypes.data_as(C.c_void_p) if fetch else None), "frontend_prepare")
        return out

    def input_points(self) -> int:
        return int(self.lib.immesh_voxelgrid_input_points(self._h) or 0)

    def step_async_raw(self, lio: "Lio", pts, leaf: float, dt=0.0, calib_laser=False, cov_gyr=0.1, cov_acc=0.1, n=None, stride=3, on_device=False):
        """The device-resident front-end chain: [calibration] -> VoxelGrid -> lio.step_async, no host round trip in between."""
        if on_device:
            ptr, cnt = C.c_void_p(int(pts)), int(n)
        else:
            a = np.ascontiguousarray(pts, dtype=np.float32)
            ptr, cnt, stride = a.ctypes.data_as(C.c_void_p), a.shape[0], a.shape[1]
        _check(self.lib, self.lib.immesh_lio_step_async_raw(lio._h, self._h, ptr, cnt, stride, 1 if on_device else 0, 1 if calib_laser else 0, C.c_float(leaf),
                                                            dt, cov_gyr, cov_acc), "lio_step_async_raw")
        lio._last_n = cnt


class Imu:
    """ImuProcess::UndistortPcl on the device (immesh_imu_*): forward propagation of the localization handle's state over the IMU
    samples of a scan + per-point motion compensation.  cfg: dict with cov_gyr, cov_acc, cov_bias_gyr, cov_bias_acc, mean_acc_norm,
    lid_R (3x3), lid_T."""

    def __init__(self, cfg: dict, max_points: int = 1 << 18, max_imu: int = 128, lib: Optional[C.CDLL] = None):
        self.lib = lib or load_library()
        c = _ImuCfg()
        for k in ("cov_gyr", "cov_acc", "cov_bias_gyr", "cov_bias_acc", "lid_T"):
            for i in range(3):
                getattr(c, k)[i] = float(np.asarray(cfg[k]).reshape(-1)[i])
        for i in range(9):
            c.lid_R[i] = float(np.asarray(cfg["lid_R"]).reshape(-1)[i])
        c.mean_acc_norm = float(cfg["mean_acc_norm"])
        c.max_points, c.max_imu = max_points, max_imu
        self._h = C.c_void_p()
        _check(self.lib, self.lib.immesh_imu_create(C.byref(c), C.byref(self._h)), "imu_create")

    def close(self):
        if self._h:
            self.lib.immesh_imu_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, last_imu7, last_lidar_end_time, last_update_time=0.0, acc_s_last=None, angvel_last=None):
        a = np.ascontiguousarray(last_imu7, dtype=np.float64)
        dp = C.POINTER(C.c_double)
        f = lambda v: None if v is None else np.ascontiguousarray(v, dtype=np.float64).ctypes.data_as(dp)  # noqa: E731
        _check(self.lib, self.lib.immesh_imu_reset(self._h, a.ctypes.data_as(dp), last_lidar_end_time, last_update_time, f(acc_s_last), f(angvel_last)), "imu_reset")

    def undistort(self, lio: "Lio", imu, pts_xyzt, lidar_beg_time):
        """imu: float64[n_imu,7] (stamp, gyr, acc); pts_xyzt: float32[n,4] (x, y, z, curvature ms).  Returns the time-sorted, compensated cloud."""
        im = np.ascontiguousarray(imu, dtype=np.float64).reshape(-1, 7)
        p = np.ascontiguousarray(pts_xyzt, dtype=np.float32).reshape(-1, 4)
        out = np.zeros_like(p)
        _check(self.lib, self.lib.immesh_imu_undistort(self._h, lio._h, im.ctypes.data_as(C.POINTER(C.c_double)), im.shape[0], p.ctypes.data_as(C.c_void_p), p.shape[0], 0,
                                                       lidar_beg_time, out.ctypes.data_as(C.c_void_p)), "imu_undistort")
        return out

    def poses(self, cap=300):
        o = np.zeros((cap, 22))
        m = self.lib.immesh_imu_get_poses(self._h, o.ctypes.data_as(C.POINTER(C.c_double)), cap)
        return o[:m].copy()


def write_ply(path: str, vertices, triangles, flips=None, lib: Optional[C.CDLL] = None):
    """save_to_ply_file layout for a snapshot (vertices float32[nv,3], triangles int32[nt,3], flips int32[nt])."""
    lib = lib or load_library()
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    t = np.ascontiguousarray(triangles, dtype=np.int32).reshape(-1, 3)
    fl = None if flips is None else np.ascontiguousarray(flips, dtype=np.int32)
    _check(lib, lib.immesh_write_ply(path.encode(), v.ctypes.data_as(C.POINTER(C.c_float)), v.shape[0], t.ctypes.data_as(C.POINTER(C.c_int32)),
                                     None if fl is None else fl.ctypes.data_as(C.POINTER(C.c_int32)), t.shape[0]), "write_ply")


def write_pcd(path: str, vertices, lib: Optional[C.CDLL] = None):
    lib = lib or load_library()
    v = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
    _check(lib, lib.immesh_write_pcd(path.encode(), v.ctypes.data_as(C.POINTER(C.c_float)), v.shape[0]), "write_pcd")


def kitti_pose_line(state, stamp: float, lib: Optional[C.CDLL] = None) -> str:
    lib = lib or load_library()
    s = np.ascontiguousarray(state, dtype=np.float64)
    buf = C.create_string_buffer(512)
    _check(lib, lib.immesh_kitti_pose_line(s.ctypes.data_as(C.POINTER(C.c_double)), stamp, buf, 512), "kitti_pose_line")
    return buf.value.decode()


def profile_enable(on: bool, lib: Optional[C.CDLL] = None):
    lib = lib or load_library()
    lib.immesh_profile_enable(int(on))   # 0 off, 1 per-kernel totals, 2 totals + launch timeline


def profile_reset(lib: Optional[C.CDLL] = None):
    (lib or load_library()).immesh_profile_reset()


def profile_report(lib: Optional[C.CDLL] = None) -> dict:
    """{kernel: (total_ms, launches)} measured with CUDA events on the launching stream."""
    lib = lib or load_library()
    n = lib.immesh_profile_report(None, 0)
    buf = C.create_string_buffer(n + 16)
    lib.immesh_profile_report(buf, n + 16)
    out = {}
    for line in buf.value.decode().splitlines():
        name, ms, cnt = line.rsplit(" ", 2)
        out[name.strip("()")] = (float(ms), int(cnt))
    return out


def profile_timeline(lib: Optional[C.CDLL] = None):
    """[(kernel, t0_ms, t1_ms)] of every launch since profile_enable(2)."""
    lib = lib or load_library()
    need = lib.immesh_profile_timeline(None, 0)
    buf = C.create_string_buffer(need + 16)
    lib.immesh_profile_timeline(buf, need + 16)
    out = []
    for line in buf.value.decode().splitlines():
        name, t0, t1 = line.rsplit(" ", 2)
        out.append((name.strip("()"), float(t0), float(t1)))
    return out


def launch_count(lib: Optional[C.CDLL] = None) -> int:
    return int((lib or load_library()).immesh_launch_count())


def graph_stats(lio: Optional["Lio"], mesh: Optional["Mesh"]) -> dict:
    """CUDA-graph replay accounting of the pipelined entry points."""
    lib = (lio or mesh).lib
    out = (C.c_int64 * 6)()
    _check(lib, lib.immesh_graph_stats(lio._h if lio else None, mesh._h if mesh else None, out), "graph_stats")
    keys = ("lio_captures", "lio_replays", "lio_failures", "mesh_captures", "mesh_replays", "mesh_failures")
    return dict(zip(keys, [int(v) for v in out]))


def host_wait_ms(lio: Optional["Lio"], mesh: Optional["Mesh"]):
    """Host milliseconds the enqueue calls spent blocked on busy staging slots since the last call: (lio, mesh)."""
    lib = (lio or mesh).lib
    out = (C.c_double * 2)()
    _check(lib, lib.immesh_host_wait_ms(lio._h if lio else None, mesh._h if mesh else None, out), "host_wait_ms")
    return float(out[0]), float(out[1])


def pipeline_mark_begin(lio: Lio):
    _check(lio.lib, lio.lib.immesh_pipeline_mark_begin(lio._h), "pipeline_mark_begin")


def pipeline_mark_end(lio: Lio, mesh: Mesh) -> float:
    ms = C.c_double(0)
    _check(lio.lib, lio.lib.immesh_pipeline_mark_end(lio._h, mesh._h, C.byref(ms)), "pipeline_mark_end")
    return ms.value


def comm_unique_id(lib: Optional[C.CDLL] = None) -> bytes:
    lib = lib or load_library()
    buf = C.create_string_buffer(128)
    _check(lib, lib.immesh_comm_unique_id(buf), "comm_unique_id")
    return buf.raw
