"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_shard_check.py
Every rank runs the same scan stream twice on its GPU: once with the VoxelMap sharded over all ranks (NCCL all-reduces
inside the C library) and once unsharded.  The sharded state must equal the unsharded one bit for bit on every rank,
and the union of the ranks' map shards must equal the unsharded map."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)
from immesh_b200 import api, synth  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = api.load_library()
    uid = [api.comm_unique_id(lib) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    cfg = api.AVIA
    n_scans = 8
    sensor, scans = synth.make_stream("avia", n_scans, seed=0, ext_T=cfg.ext_T)
    handles = {"sharded": api.Lio(cfg, lib=lib), "single": api.Lio(cfg, lib=lib)}
    handles["sharded"].shard(rank, world, uid[0])
    for h in handles.values():
        h.set_pose(scans[0]["R_true"], scans[0]["t_true"])
        s = h.get_state()
        s[12:15] = (scans[1]["t_true"] - scans[0]["t_true"]) / scans[0]["dt"]
        h.set_state(s)
        h.voxel_map_init(scans[0]["body_full"])
    ok = True
    for k in range(1, n_scans):
        ss, _ = handles["sharded"].step(scans[k]["body_ds"], scans[k]["dt"])
        s1, _ = handles["single"].step(scans[k]["body_ds"], scans[k]["dt"])
        if not np.array_equal(ss, s1):
            ok = False
            print(f"[rank {rank}] scan {k}: sharded state differs, max |d| = {np.abs(ss - s1).max()}", flush=True)
    dumps = [None] * world
    dist.all_gather_object(dumps, handles["sharded"].dump_map())
    ref = handles["single"].dump_map()

    def blocks(d):
        out, start = {}, 0
        for i in range(1, len(d) + 1):
            if i == len(d) or d[i, 3] == 0:
                out[tuple(d[start, :3])] = d[start:i]
                start = i
        return out
    merged = {}
    for d in dumps:
        merged.update(blocks(d))
    br = blocks(ref)
    map_ok = merged.keys() == br.keys() and all(np.array_equal(merged[k], br[k]) for k in br)
    sizes = [len(d) for d in dumps]
    if rank == 0:
        print(f"ranks={world} state_bit_exact={ok} map_union_bit_exact={map_ok} shard_rows={sizes} single_rows={len(ref)}", flush=True)
    flag = torch.tensor([int(ok and map_ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
