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
