"""CPU tier, world_size 2 over gloo: the sharded-VoxelMap localization path (root voxels owned by key hash, replicated scan,
two integer all-reduces per IESKF iteration) executed by the host emulation of the kernel bodies must reproduce the
single-process oracle bit for bit: identical state on both ranks, and the union of the two map shards == the oracle map."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_scans, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from immesh_b200 import api, build, synth
    lib = api.load_library(build.EMU)
    lib.emu_shard_bits.restype = C.POINTER(C.c_uint32)
    lib.emu_shard_acc.restype = C.POINTER(C.c_uint64)
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", n_scans, seed=0, ext_T=cfg.ext_T)
    g = api.Lio(cfg, lib=lib)
    lib.emu_lio_set_shard(g._h, rank, world)
    g.set_pose(scans[0]["R_true"], scans[0]["t_true"])
    s = g.get_state()
    s[12:15] = (scans[1]["t_true"] - scans[0]["t_true"]) / scans[0]["dt"]
    g.set_state(s)
    g.voxel_map_init(scans[0]["body_full"])           # every rank keeps only the root voxels it owns
    states = []
    for k in range(1, n_scans):
        body = np.ascontiguousarray(scans[k]["body_ds"], dtype=np.float32)
        g.predict(scans[k]["dt"])
        lib.emu_shard_begin(g._h, body.ctypes.data_as(C.POINTER(C.c_float)), body.shape[0])
        for it in range(cfg.max_iteration):
            lib.emu_shard_pass1(g._h)
            nw = C.c_int(0)
            bp = lib.emu_shard_bits(g._h, C.byref(nw))
            bits = np.ctypeslib.as_array(bp, shape=(nw.value,))
            t = torch.from_numpy(bits.astype(np.int64))       # disjoint bits: SUM == OR
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            bits[:] = t.numpy().astype(np.uint32)
            lib.emu_shard_pass2(g._h, it)
            ap = lib.emu_shard_acc(g._h, it)
            acc = np.ctypeslib.as_array(ap, shape=(60,))
            t = torch.from_numpy(acc.astype(np.int64))        # two's complement sums
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            acc[:] = t.numpy().astype(np.uint64)
            if lib.emu_shard_solve(g._h, it):
                break
        g._last_n = body.shape[0]
        g.map_incremental_grow()
        states.append(g.get_state())
    np.save(os.path.join(out_dir, f"states_{rank}.npy"), np.array(states))
    np.save(os.path.join(out_dir, f"map_{rank}.npy"), g.dump_map())
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_lio_two_ranks_equals_oracle(tmp_path, built):
    from immesh_b200 import api, synth
    from oracle_api import OracleLio
    n_scans = 5
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_scans, str(tmp_path)), nprocs=2, join=True)
    cfg = api.AVIA
    sensor, scans = synth.make_stream("avia", n_scans, seed=0, ext_T=cfg.ext_T)
    o = OracleLio(cfg, sum_mode=0)
    o.set_pose(scans[0]["R_true"], scans[0]["t_true"])
    s = o.get_state()
    s[12:15] = (scans[1]["t_true"] - scans[0]["t_true"]) / scans[0]["dt"]
    o.set_state(s)
    o.voxel_map_init(scans[0]["body_full"])
    ref_states = []
    for k in range(1, n_scans):
        o.predict(scans[k]["dt"])
        o.lio_state_estimation(scans[k]["body_ds"])
        o.map_incremental_grow(scans[k]["body_ds"])
        ref_states.append(o.get_state())
    ref_states = np.array(ref_states)
    s0 = np.load(tmp_path / "states_0.npy")
    s1 = np.load(tmp_path / "states_1.npy")
    assert np.array_equal(s0, s1), "ranks disagree on the state"
    assert np.array_equal(s0, ref_states), "sharded state differs from the single-process oracle"
    m = np.concatenate([np.load(tmp_path / "map_0.npy"), np.load(tmp_path / "map_1.npy")])
    ref = o.dump_map()
    assert m.shape == ref.shape
    # merge the shards: rows of one root voxel are contiguous in each dump; sort voxel blocks by key
    def blocks(d):
        out, start = {}, 0
        for i in range(1, len(d) + 1):
            if i == len(d) or d[i, 3] == 0:      # path 0 = a root row starts a new block
                out[tuple(d[start, :3])] = d[start:i]
                start = i
        return out
    bm, br = blocks(m), blocks(ref)
    assert bm.keys() == br.keys()
    for key in br:
        assert np.array_equal(bm[key], br[key]), key
    n0 = len(np.load(tmp_path / "map_0.npy"))
    assert 0.3 < n0 / len(ref) < 0.7               # the shards are balanced


def _gather_padded(arr, world):
    """all_gather of variable-length uint8 buffers: lengths first, then buffers padded to the maximum."""
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    n = torch.tensor([raw.size], dtype=torch.int64)
    lens = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(lens, n)
    m = max(int(x) for x in lens)
    buf = torch.zeros(max(m, 1), dtype=torch.uint8)
    buf[:raw.size] = torch.from_numpy(raw.copy())
    outs = [torch.zeros(max(m, 1), dtype=torch.uint8) for _ in range(world)]
    dist.all_gather(outs, buf)
    return [o.numpy()[:int(l)].copy() for o, l in zip(outs, lens)]


def _mesh_worker(rank, world, port, n_frames, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from immesh_b200 import api, build
    from mesh_common import SMALL, world_scans
    lib = api.load_library(build.EMU)
    u64p, i32p = C.POINTER(C.c_uint64), C.POINTER(C.c_int)
    cfg = api.MeshConfig(**SMALL)
    g = api.Mesh(cfg, lib=lib)
    lib.emu_mesh_set_shard(g._h, rank, world)
    owned = []
    for k, (w, t) in enumerate(world_scans("avia", n_frames, 3, 20000)):
        w = np.ascontiguousarray(w, dtype=np.float32)
        t = np.ascontiguousarray(t, dtype=np.float64)
        assert lib.emu_mesh_phase_a(g._h, w.ctypes.data_as(C.POINTER(C.c_float)), w.shape[0], t.ctypes.data_as(C.POINTER(C.c_double))) == 0
        c = (C.c_int * 3)()
        lib.emu_mesh_xcounts(g._h, c)
        sm = np.zeros(c[0] * 32, dtype=np.uint8)
        lib.emu_mesh_xget(g._h, sm.ctypes.data_as(C.c_void_p), None, None, None)
        for r, buf in enumerate(_gather_padded(sm, world)):          # exchange 1: smoothed positions
            if r != rank and buf.size:
                lib.emu_mesh_xapply_smooth(g._h, buf.ctypes.data_as(C.c_void_p), buf.size // 32)
        lib.emu_mesh_phase_b(g._h)
        lib.emu_mesh_xcounts(g._h, c)
        face = np.zeros((c[1], 4), dtype=np.int32); word = np.zeros(c[1], dtype=np.uint64); rem = np.zeros((c[2], 4), dtype=np.int32)
        lib.emu_mesh_xget(g._h, None, face.ctypes.data_as(i32p), word.ctypes.data_as(u64p), rem.ctypes.data_as(i32p))
        owned.append(int(c[1]))
        faces, words, rems = _gather_padded(face, world), _gather_padded(word, world), _gather_padded(rem, world)
        for r in range(world):                                          # exchange 2: facets + removals of every rank
            f = np.ascontiguousarray(faces[r].view(np.int32)); wd = np.ascontiguousarray(words[r].view(np.uint64)); rm = np.ascontiguousarray(rems[r].view(np.int32))
            lib.emu_mesh_xapply_lists(g._h, f.ctypes.data_as(i32p), wd.ctypes.data_as(u64p), wd.size, rm.ctypes.data_as(i32p), rm.size // 4)
        assert lib.emu_mesh_phase_c(g._h) == 0
    v, tri, fl = g.snapshot()
    np.savez(os.path.join(out_dir, f"mesh_{rank}.npz"), v=v, tri=tri, fl=fl, owned=np.array(owned))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_mesher_two_ranks_equals_oracle(tmp_path, built):
    """Per-voxel meshing stage sharded by voxel owner over 2 ranks (replicated append, two list exchanges per frame):
    both replicas must hold the oracle's vertex array, facet set and orientation flags."""
    from immesh_b200 import api
    from mesh_common import SMALL, world_scans
    from oracle_api import OracleMesh
    n_frames = 5
    port = _free_port()
    mp.spawn(_mesh_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    o = OracleMesh(api.MeshConfig(**SMALL))
    for k, (w, t) in enumerate(world_scans("avia", n_frames, 3, 20000)):
        o.push_frame(w, t, k)
    vo, to, fo = o.snapshot()
    r0, r1 = np.load(tmp_path / "mesh_0.npz"), np.load(tmp_path / "mesh_1.npz")
    for r in (r0, r1):
        assert np.array_equal(r["v"], vo)
        assert r["tri"].shape == to.shape and np.array_equal(r["tri"], to)
        assert np.array_equal(r["fl"], fo)
    assert len(to) > 1000
    share = r0["owned"].sum() / max(1, r0["owned"].sum() + r1["owned"].sum())
    assert 0.3 < share < 0.7, share               # both ranks meshed a comparable number of voxels
