"""Multi-GPU parity check (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_shard_check.py
Every rank runs the same scan stream twice on its GPU: once with the VoxelMap sharded over all ranks (exchanges fused into
the kernels over NVLink peer windows; IMMESH_SHARD_NCCL=1 selects the NCCL all-reduce / all-gather baseline) and once unsharded.  The sharded state must equal the unsharded one bit for bit on every rank,
and the union of the ranks' map shards must equal the unsharded map.  The mesher runs sharded (per-voxel stage by voxel
owner, two all-gathers per frame) next to an unsharded instance: vertices, facet set and flags must be identical."""
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
    uid = [api.comm_unique_id(lib) if rank == 0 else None, api.comm_unique_id(lib) if rank == 0 else None]
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
    # the mesher: per-voxel stage sharded by voxel owner (own NCCL communicator) vs one GPU, fed by the same poses
    from mesh_common import SMALL
    meshes = {"sharded": api.Mesh(api.MeshConfig(**SMALL), lib=lib), "single": api.Mesh(api.MeshConfig(**SMALL), lib=lib)}
    meshes["sharded"].shard(rank, world, uid[1])
    ok = True
    mesh_ok = True
    for k in range(1, n_scans):
        ss, _ = handles["sharded"].step(scans[k]["body_ds"], scans[k]["dt"])
        s1, _ = handles["single"].step(scans[k]["body_ds"], scans[k]["dt"])
        if not np.array_equal(ss, s1):
            ok = False
            print(f"[rank {rank}] scan {k}: sharded state differs, max |d| = {np.abs(ss - s1).max()}", flush=True)
        meshes["sharded"].push_frame_from_lio(handles["sharded"], scans[k]["body_full"])
        meshes["single"].push_frame_from_lio(handles["single"], scans[k]["body_full"])
        (va, ta, fa), (vb, tb, fb) = meshes["sharded"].snapshot(), meshes["single"].snapshot()
        same = np.array_equal(va, vb) and ta.shape == tb.shape and np.array_equal(ta, tb) and np.array_equal(fa, fb)
        if not same:
            mesh_ok = False
            print(f"[rank {rank}] frame {k}: sharded mesh differs: vertices {va.shape} vs {vb.shape}, facets {ta.shape} vs {tb.shape}", flush=True)
    n_facets = len(meshes["single"].snapshot()[1])
    owned = meshes["sharded"].work_stats()["voxels_meshed"]
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
        print(f"ranks={world} state_bit_exact={ok} map_union_bit_exact={map_ok} shard_rows={sizes} single_rows={len(ref)} "
              f"mesh_replicas_bit_exact={mesh_ok} facets={n_facets} voxels_meshed_last_frame_rank0={owned} "
              f"transport_voxelmap={handles['sharded'].shard_transport()} transport_mesher={meshes['sharded'].shard_transport()}", flush=True)
    flag = torch.tensor([int(ok and map_ok and mesh_ok)], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
