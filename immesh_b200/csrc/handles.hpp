// immesh_b200 -- handle definitions shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>

#include <vector>

#include "common_host.hpp"
#include "lio_core.cuh"
#include "mesh_voxel.cuh"
#include "peer_win.cuh"

using immesh::LioParams; using immesh::VoxelMapDev; using immesh::ScanBuf; using immesh::LioCtrl;
using immesh::MeshParams; using immesh::MeshDev; using immesh::FrameBuf; using immesh::FramePose;

// staging depth of the pipelined entry points: the host may queue this many scans / frames ahead of the device before it has to wait
// (2 was enough for the device, but a host hiccup of ~1 ms then drained the pipeline: seen as single ranks of an 8-rank run falling
// back to the blocking-call rate)
#define IM_SLOTS 4

struct immesh_lio {
    LioParams P;
    VoxelMapDev map;
    ScanBuf sb;
    LioCtrl* d_ctrl = nullptr;
    int* d_counters = nullptr;  // node_count, chunk_bump, avail_top, pending_n, err, n_roots, n_touched, seg_top, work_counter
    int* d_sorted = nullptr;
    int* d_complex = nullptr;   // touched root voxels left to the general warp-per-voxel kernel (the others are finished by k_grow_simple)
    double* d_ptpl = nullptr;
    float* d_body_own = nullptr;  // scan buffers owned by the handle, 2 slots of max_scan*3 (sb.body points into them unless the caller passed a device pointer)
    cudaStream_t stream_up = nullptr;   // upload stream: the H2D copy of scan k+1 overlaps the kernels of scan k
    cudaEvent_t ev_up[IM_SLOTS] = {};
    float* h_body = nullptr;   // pinned staging, IM_SLOTS slots
    double* h_state = nullptr; // pinned, 2 slots of (IM_STATE_DOUBLES + 64)
    cudaEvent_t ev_slot[IM_SLOTS] = {};  // completion of the step that used staging slot s
    int slot_busy[IM_SLOTS] = {};
    int step_counter = 0;
    immesh::ScanDyn* h_dyn = nullptr;   // pinned, 2 slots: per-scan inputs of the launch sequence (one H2D per scan)
    immesh::ScanDyn dyn_last = {};         // host copy of the block last uploaded
    immesh::LioOut* h_out = nullptr;    // pinned, 2 slots: what a scan returns to the host (one D2H per scan)
    const int* next_n_dev = nullptr;    // device-resident point count for the next upload (immesh_lio_step_async_dev_n)
    int scan_counter = 0;               // scans queued through the step entry points; pose ring slot = scan_counter & 7
    int pose_pub_idx = -1;              // scan index whose converged pose is in LioCtrl::pose_ring (-1: not published)
    int timed_last = 0;                 // the last step recorded its stage timing events
    double host_wait_ms = 0;            // host time spent blocked on a busy staging slot inside the enqueue calls (back-pressure, not work)
    cudaEvent_t ev_pose = nullptr;      // "pose of the last scan published" (the mesher's stream waits on it)
    cudaEvent_t ev_mark = nullptr;   // pipeline timing mark (begin)
    void* nccl_comm = nullptr;       // ncclComm_t when the VoxelMap is sharded over several GPUs
    unsigned int* d_bits = nullptr;  // [2][words] exists / matched-in-own-voxel bit words of the sharded residual pass
    immesh::PeerWindow win;          // peer window of the sharded residual pass (NVLink peer memory, replaces the NCCL all-reduces)
    int words_cap = 0;
    int* h_ints = nullptr;     // pinned
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t cap = 0;
    int max_nodes = 0, max_chunks = 0, max_scan = 0;
    int n_sm = 148;
    int last_n = 0;
    immesh::GraphCtx graph;   // replay of the per-scan launch sequence (pipelined API)
    int use_graph = 1;
    int bps = 4;              // (mesh: 3 by default, see immesh_mesh_create) resident blocks per SM of the persistent per-voxel kernels (headroom for the other stream)
    double last_ms[3] = {0, 0, 0};
    std::vector<void*> allocs;
};

struct immesh_mesh {
    MeshParams P;
    MeshDev M;
    FrameBuf F;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    float* d_pts = nullptr;   // 2 slots of max_frame_points*3
    float* d_body = nullptr;  // 2 slots, staging for body-frame scans handed over by the localization handle
    float* h_pts = nullptr;   // pinned, 2 slots
    int* h_cnt = nullptr;     // pinned, 2 slots of 32
    FramePose* d_fp = nullptr;  // 2 slots
    FramePose* h_fp = nullptr;  // pinned, 2 slots
    immesh::FrameDyn* h_dyn = nullptr;  // pinned, 2 slots: per-frame inputs of the launch sequence
    immesh::FrameDyn* d_dyn = nullptr;  // device copy read by every kernel of the frame (F.dyn)
    int timed_last = 0;
    double host_wait_ms = 0;
    cudaEvent_t ev_in[IM_SLOTS] = {};    // inputs of slot s ready (recorded on the producer stream)
    cudaEvent_t ev_done[IM_SLOTS] = {};  // frame of slot s finished (mesh stream)
    int inflight[IM_SLOTS] = {};
    cudaEvent_t ev_mark = nullptr, ev_sync = nullptr;  // pipeline timing mark (end) / cross-stream join
    cudaStream_t stream2 = nullptr;   // side stream: warp-level triangulation, concurrent with the block-level one
    cudaStream_t stream3 = nullptr;   // side stream: pull (incidence-list walk), concurrent with the triangulation
    cudaStream_t stream_up = nullptr; // upload stream: the H2D copy of frame k+1 overlaps the kernels of frame k
    cudaEvent_t ev_up[IM_SLOTS] = {};
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_join3 = nullptr;
    int pending_rc = 0;
    void* nccl_comm = nullptr;        // ncclComm_t when the per-voxel stage is sharded over several GPUs
    unsigned char *d_seg1 = nullptr, *d_seg2 = nullptr, *d_recv1 = nullptr, *d_recv2 = nullptr;   // exchange segments
    immesh::PeerWindow win;           // peer window: [flags][recv1 x n][recv2 x n] when the segments are pushed over NVLink
    int* d_xdone = nullptr;           // block-done counters of the two push kernels
    size_t seg1_bytes = 0, seg2_bytes = 0;
    int* d_snap_tri = nullptr;
    int* d_snap_flip = nullptr;
    int* d_snap_n = nullptr;
    int max_frame_points = 0;
    immesh::GraphCtx graph;   // replay of the per-frame launch sequence (pipelined API)
    int use_graph = 1;
    int bps = 4;              // (mesh: 3 by default, see immesh_mesh_create) resident blocks per SM of the persistent per-voxel kernels (headroom for the other stream)
    int frame_counter = 0;
    int dilate_bps = 3;       // resident blocks per SM of the dilation kernel
    int warp_nmax = 96;       // largest dilated set triangulated by a single warp
    int n_sm = 148;
    size_t ccap = 0;
    double last_ms[4] = {0, 0, 0, 0};
    int last_cnt[32];
    std::vector<void*> allocs;
};
