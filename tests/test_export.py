"""CPU tier: immesh_write_ply (host-side export of a snapshot, save_to_ply_file layout) -- header, vertex block and the reference's
face orientation rule (m_index_flip != 0 keeps p0 p1 p2, == 0 swaps the last two: mesh_rec_geometry.cpp:108-121)."""
import numpy as np

from immesh_b200 import api


def test_write_ply_roundtrip(tmp_path):
    lib = api.load_library()   # host-only entry point: no CUDA call is made
    rng = np.random.default_rng(0)
    v = rng.normal(0, 3, (50, 3)).astype(np.float32)
    t = np.sort(rng.integers(0, 50, (80, 3)), axis=1).astype(np.int32)
    fl = rng.integers(0, 2, 80).astype(np.int32)
    path = str(tmp_path / "mesh.ply")
    api.write_ply(path, v, t, fl, lib=lib)
    raw = open(path, "rb").read()
    head, body = raw.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    assert lines[0] == "ply" and lines[1] == "format binary_little_endian 1.0"
    assert "element vertex 50" in lines and "element face 80" in lines and "property list uchar int vertex_indices" in lines
    vb = np.frombuffer(body[:50 * 12], dtype="<f4").reshape(50, 3)
    assert np.array_equal(vb, v)
    faces = np.frombuffer(body[50 * 12:], dtype=np.dtype([("n", "u1"), ("i", "<i4", 3)]))
    assert len(faces) == 80 and np.all(faces["n"] == 3)
    want = np.where(fl[:, None] != 0, t, t[:, [0, 2, 1]])
    assert np.array_equal(faces["i"], want)
    # a face pointing outside the vertex array is refused
    bad = t.copy()
    bad[3, 1] = 50
    try:
        api.write_ply(path, v, bad, fl, lib=lib)
        assert False
    except RuntimeError:
        pass


def test_write_pcd_layout(tmp_path):
    lib = api.load_library()
    rng = np.random.default_rng(1)
    v = rng.normal(0, 5, (77, 3)).astype(np.float32)
    path = str(tmp_path / "mesh.ply.pcd")
    api.write_pcd(path, v, lib=lib)
    raw = open(path, "rb").read()
    head, body = raw.split(b"DATA binary\n", 1)
    lines = head.decode().splitlines()
    assert lines[0].startswith("# .PCD v0.7") and "FIELDS x y z" in lines and "SIZE 4 4 4" in lines and "TYPE F F F" in lines
    assert "WIDTH 77" in lines and "HEIGHT 1" in lines and "POINTS 77" in lines
    assert np.array_equal(np.frombuffer(body, dtype="<f4").reshape(-1, 3), v)


def test_kitti_pose_line_matches_numpy():
    """Voxel_mapping::kitti_log (voxel_mapping_common.cpp:43-70): T_cam = T_l2c [R t] T_l2c^-1 and its quaternion, against numpy / scipy."""
    from scipy.spatial.transform import Rotation
    lib = api.load_library()
    L = np.array([[0.00554604, -0.999971, -0.00523653, 0.0316362], [-0.000379382, 0.00523451, -0.999986, 0.0380934],
                  [0.999985, 0.00554795, -0.000350341, 0.409066], [0, 0, 0, 1.0]])
    rng = np.random.default_rng(2)
    for _ in range(20):
        R = Rotation.from_rotvec(rng.normal(0, 1.5, 3)).as_matrix()
        t = rng.normal(0, 50, 3)
        s = np.zeros(348)
        s[:9] = R.reshape(9)
        s[9:12] = t
        line = api.kitti_pose_line(s, 12.5, lib=lib)
        f = [float(x) for x in line.split()]
        assert len(f) == 8 and line.endswith("\n") and f[0] == 12.5
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        B = L @ T @ np.linalg.inv(L)
        assert np.allclose(f[1:4], B[:3, 3], atol=2e-6)
        q = Rotation.from_matrix(B[:3, :3]).as_quat()     # x y z w
        got = np.array(f[4:8])
        assert min(np.abs(got - q).max(), np.abs(got + q).max()) < 2e-5     # %lf = 6 decimals; T_l2c is orthonormal to ~1e-6 only
