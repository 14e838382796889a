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
