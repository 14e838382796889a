"""Where does the pipelined scan time go?  Runs the stream of bench.py in a few reduced configurations (GPU box only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from immesh_b200 import api

K = 20
cfg, sensor, scans = bench.get_stream(1 + 8 + 6 * K + 8)
lib = api.load_library()
lio = api.Lio(cfg, lib=lib)
mesh = api.Mesh(api.MeshConfig(), lib=lib)
lio.set_state(bench.init_state_vec(scans))
lio.voxel_map_init(scans[0]["body_full"])
dev = torch.device("cuda", 0)
d_ds = [torch.from_numpy(s["body_ds"]).to(dev) for s in scans]
d_full = [torch.from_numpy(s["body_full"]).to(dev) for s in scans]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
k = 1
for _ in range(8):
    lio.step_dev(d_ds[k].data_ptr(), d_ds[k].shape[0], scans[k]["dt"])
    mesh.push_frame_from_lio(lio, d_full[k].data_ptr(), d_full[k].shape[0], on_device=True)
    k += 1

def run(name, do_flush, do_mesh, n=K):
    global k
    torch.cuda.synchronize()
    api.pipeline_mark_begin(lio)
    t0 = time.perf_counter()
    first2 = None
    for i in range(n):
        if do_flush:
            lio.enqueue_memset(flush.data_ptr(), flush.numel())
        lio.step_async(d_ds[k].data_ptr(), d_ds[k].shape[0], scans[k]["dt"], on_device=True)
        if do_mesh:
            mesh.push_frame_from_lio_async(lio, d_full[k].data_ptr(), d_full[k].shape[0], on_device=True)
        k += 1
        if i == 1:
            first2 = (time.perf_counter() - t0) * 1e3 / 2
    ms = api.pipeline_mark_end(lio, mesh) / n
    lio.wait(); mesh.wait()
    print(f"{name:34s} {ms:.4f} ms/scan   host enqueue of the first 2 scans (unthrottled): {first2:.4f} ms/scan")

run("lio only, no flush", False, False)
run("lio only, flush", True, False)
run("lio+mesh, no flush", False, True)
run("lio+mesh, flush (bench)", True, True)
print(api.graph_stats(lio, mesh))

# ---- timeline of three pipelined scans in steady state (direct launches with events around every kernel)
if os.environ.get("IMMESH_TIMELINE", "1") == "1":
    api.profile_reset(lib)
    api.profile_enable(2, lib)
    n = 6
    for i in range(n):
        lio.enqueue_memset(flush.data_ptr(), flush.numel())
        lio.step_async(d_ds[k].data_ptr(), d_ds[k].shape[0], scans[k]["dt"], on_device=True)
        mesh.push_frame_from_lio_async(lio, d_full[k].data_ptr(), d_full[k].shape[0], on_device=True)
        k += 1
    lio.wait(); mesh.wait()
    torch.cuda.synchronize()
    lio.wait(); mesh.wait()
    tl = sorted(api.profile_timeline(lib), key=lambda r: r[1])
    api.profile_enable(0, lib)
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/timeline.txt", "w") as f:
        t_base = tl[0][1]
        for name, t0, t1 in tl:
            f.write(f"{(t0 - t_base) * 1e3:9.1f} {(t1 - t_base) * 1e3:9.1f} {(t1 - t0) * 1e3:7.1f}  {name}\n")
    print("timeline spans:", len(tl))
