// P3P instance collection + batched minimal-solver pose sampling (device-resident).
//
// Replaces, for one camera of the window:
//   reference gpu-kernels/collect_p3p_instances.cu:57-250   (instance maps)
//   reference voldor/geometry.cpp:70-80                      (host raster-order compaction)
//   reference gpu-kernels/solve_batch_lambdatwist.cu:11-102  (8192 P4P hypotheses, cuRAND indices)
//   reference gpu-kernels/solve_batch_ap3p.cu:331-437        (alternative AP3P solver)
//   reference voldor/geometry.cpp:156-165,191                (finite filter of hypotheses, rvec x scale)
// with everything kept on the device: maps -> order-preserving compaction -> hypotheses -> pose pool.
#pragma once
#include "common.cuh"

namespace vb {

struct CollectParams {
    int active_idx;
    float rigidness_thresh, rigidness_sum_thresh, sample_min_depth, sample_max_depth;
    int max_trace_on_flow;
};

struct Collector {
    int w = 0, h = 0;
    TexStack<float2> flows_own;
    Plane<float> rig_own, depth_own;
    // views actually used by the kernels (own storage for the ABI path, DepthEM's for the window pipeline)
    const TexStack<float2>* flows = nullptr;
    const float* rig = nullptr;
    int rig_pitch = 0;
    size_t rig_plane = 0;
    const float* depth = nullptr;
    int depth_pitch = 0;
    CamBlock cam;
    const CamBlock* d_cam = nullptr;  // when set, the compacting kernel reads the poses from device memory instead

    float* p2_map = nullptr;  // [h*w][2]  NaN where invalid
    float* p3_map = nullptr;  // [h*w][3]
    float* p2c = nullptr;     // compacted instances (raster order), capacity w*h+1
    float* p3c = nullptr;
    int* d_count = nullptr;  // number of compacted instances
    // single-pass compaction state (decoupled look-back): tile descriptors, ticket counter, launch epoch
    unsigned long long* scan_desc = nullptr;
    unsigned int* scan_ticket = nullptr;
    unsigned int ticket_total = 0, scan_epoch = 0;
    int map_capacity = 0;
    cudaStream_t stream = nullptr;

    int ensure(int w_, int h_, int N);
    void use_own_views();
    // instance maps for camera P.active_idx; with `compact` also fills p2c/p3c/d_count.  Asynchronous.
    int collect(int N, const CollectParams& P, bool compact);
};

// Per-hypothesis uniform draws of the reference sampler: 4 values of curand_uniform() from
// XORWOW(seed 233, subsequence idx, offset 0) — constants, because the reference re-seeds on every call
// (solve_batch_lambdatwist.cu:44-48,81).  Tabulated once per capacity.
// The table is a constant of (seed, index), so ONE table serves every execution context of the process: ensure() is
// serialised, and a table that was handed out is never freed (another context's kernels may still be reading it).
struct HypothesisDraws {
    // returns the table (>= n_poses entries) or nullptr on a CUDA error
    const float4* ensure(int n_poses, cudaStream_t s);
};
HypothesisDraws& global_draws();

// Batched minimal solver.  d_n_pts may be null (then n_pts_host is used).
int solve_batch_p3p_device(const float* d_p3s, const float* d_p2s, const int* d_n_pts, int n_pts_host,
                           float fx, float fy, float cx, float cy, float* d_rvecs, float* d_tvecs, int n_poses,
                           bool use_ap3p, cudaStream_t s);

// Keep finite hypotheses in hypothesis order; rvec is multiplied by rvec_scale on the way out.
// pool: [n_poses][6], d_used: count.
int filter_pose_pool(const float* d_rvecs, const float* d_tvecs, int n_poses, float rvec_scale, float* d_pool,
                     int* d_used, cudaStream_t s);

}  // namespace vb
