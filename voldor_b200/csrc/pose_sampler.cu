// P3P instance collection, on-device compaction and batched pose hypotheses for sm_100a.
// See pose_sampler.cuh for the reference call sites this replaces.
#include "pose_sampler.cuh"
#include "geometry.cuh"
#include "p3p_twist_quad.cuh"
#include "p3p_ap3p_quad.cuh"
#include "rotation.cuh"
#include <curand_kernel.h>
#include <algorithm>
#include <cstdlib>
#include <mutex>

namespace vb {

namespace {

constexpr int kBlock = 256;
constexpr int kCompactTile = 1024;  // pixels per tile of the single-pass compaction (shorter look-back chain)

struct CollectView {
    int N, w, h;
    cudaTextureObject_t flows_tex;
    const float* rig;
    int rig_pitch;
    size_t rig_plane;
    const float* depth;
    int depth_pitch;
    float* p2_map;
    float* p3_map;
};

__device__ __forceinline__ float quiet_nan() { return __int_as_float(0x7fffffff); }  // CUDART_NAN_F

// the P3P instance of one pixel, NaN where invalid (reference: collect_p3p_instances.cu:70-145)
struct Instance {
    float p2x, p2y, p3x, p3y, p3z;
};
__device__ __forceinline__ Instance pixel_instance(const CollectView& A, const CamBlock& C, const CollectParams& P,
                                                   int i) {
    float p2x = quiet_nan(), p2y = quiet_nan();
    float p3x = quiet_nan(), p3y = quiet_nan(), p3z = quiet_nan();
    const int x = i % A.w, y = i / A.w;
    const float depth = A.depth[(size_t)y * A.depth_pitch + x];
    const float* rig_px = A.rig + (size_t)y * A.rig_pitch + x;
    bool ok = !(depth < P.sample_min_depth || (P.sample_max_depth > 0 && depth > P.sample_max_depth));
    if (ok && P.rigidness_sum_thresh > (float)(A.N + 1)) {
        // only reachable for thresholds above N+1; sequential sum as in the reference
        float sum = 0;
        for (int f = 0; f < A.N; f++) sum = f_add(sum, rig_px[(size_t)f * A.rig_plane]);
        if (sum < P.rigidness_sum_thresh) ok = false;
    }
    int n_trace = 0;
    if (ok) {
        float trace_product = 1;
        const int last = P.max_trace_on_flow > 0 ? max(0, P.active_idx - P.max_trace_on_flow + 1) : 0;
        for (int f = P.active_idx; f >= last; f--) {
            trace_product = f_mul(trace_product, rig_px[(size_t)f * A.rig_plane]);
            if (trace_product > P.rigidness_thresh)
                n_trace++;
            else
                break;
        }
        if (n_trace <= 0) ok = false;
    }
    if (ok) {
        const float fw = (float)A.w, fh = (float)A.h;
        bool out_of_view = false;
        float px = 0, py = 0, ox, oy, oz;
        backproject(C, (float)x, (float)y, depth, ox, oy, oz);
        const int first_traced = P.active_idx - n_trace + 1;
        for (int f = 0; f <= P.active_idx; f++) {
            if (f >= first_traced) {
                if (f == first_traced) project(C, ox, oy, oz, px, py);
                if (px > 0 && px < fw && py > 0 && py < fh) {  // strict > 0 here
                    const float2 d2 = fetch_stack<float2>(A.flows_tex, px, py, f, A.h);
                    px = f_add(px, d2.x);
                    py = f_add(py, d2.y);
                } else {
                    out_of_view = true;
                    break;
                }
            }
            if (f < P.active_idx) rigid_move(C.R[f], C.t[f], ox, oy, oz);
        }
        if (!out_of_view && oz > P.sample_min_depth && (P.sample_max_depth <= 0 || oz < P.sample_max_depth)) {
            p2x = px, p2y = py;
            p3x = ox, p3y = oy, p3z = oz;
        }
    }
    return Instance{p2x, p2y, p3x, p3y, p3z};
}
// validity rule of the host compaction loop (reference: voldor/geometry.cpp:72)
__device__ __forceinline__ bool instance_valid(const Instance& v) {
    return isfinite(f_add(f_add(f_add(f_add(v.p2x, v.p2y), v.p3x), v.p3y), v.p3z));
}

// one thread per pixel in raster order: dense NaN-filled maps (the ABI's output layout)
__global__ void __launch_bounds__(kBlock)
    k_collect(const CollectView A, const __grid_constant__ CamBlock C, const CollectParams P) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    const int npx = A.w * A.h;
    Instance v{quiet_nan(), quiet_nan(), quiet_nan(), quiet_nan(), quiet_nan()};
    if (i < npx) {
        v = pixel_instance(A, C, P, i);
        A.p2_map[2 * (size_t)i] = v.p2x;
        A.p2_map[2 * (size_t)i + 1] = v.p2y;
        A.p3_map[3 * (size_t)i] = v.p3x;
        A.p3_map[3 * (size_t)i + 1] = v.p3y;
        A.p3_map[3 * (size_t)i + 2] = v.p3z;
    }
}

// ------------------------------------------------------------------------------------------------
// Single-pass collection + raster-order compaction (window pipeline: the dense maps are never needed there).
// Replaces the reference's map kernel + D2H + host loop (voldor/geometry.cpp:70-80) with one launch: blocks take
// raster tiles in ticket order and obtain their output offset by decoupled look-back over the descriptors
// {epoch | status | count} published by their predecessors, so the instance order is exactly the raster order.
// Descriptors carry a launch epoch, so nothing has to be cleared between launches.
// ------------------------------------------------------------------------------------------------
struct ScanState {
    unsigned long long* desc;  // one per tile
    unsigned int* ticket;      // running tile counter
    unsigned int ticket_base;  // value of *ticket when this launch starts
    unsigned int epoch;        // 1.. , never 0
};
__device__ __forceinline__ unsigned long long scan_pack(unsigned epoch, unsigned status, unsigned value) {
    return ((unsigned long long)epoch << 34) | ((unsigned long long)status << 32) | value;
}
enum { SCAN_AGGREGATE = 1, SCAN_INCLUSIVE = 2 };

template <int TILE, bool CAM_IN_MEMORY>
__global__ void __launch_bounds__(TILE)
    k_collect_compact(const CollectView A, const __grid_constant__ CamBlock Cparam, const CamBlock* __restrict__ d_cam,
                      const CollectParams P, ScanState S, float* p2c, float* p3c, int* total) {
    __shared__ unsigned s_tile;
    __shared__ int warp_counts[TILE / 32];
    __shared__ int s_prefix;
    // poses either by value (kernel parameter) or from device memory: inside an EM iteration the mean-shift kernel
    // of camera i-1 writes the pose this launch needs, without a host round trip
    __shared__ CamBlock s_cam;
    if (CAM_IN_MEMORY) {
        for (int k = threadIdx.x; k < (int)(sizeof(CamBlock) / sizeof(float)); k += TILE)
            reinterpret_cast<float*>(&s_cam)[k] = reinterpret_cast<const float*>(d_cam)[k];
    }
    if (threadIdx.x == 0) s_tile = atomicAdd(S.ticket, 1u) - S.ticket_base;
    __syncthreads();
    const CamBlock& C = CAM_IN_MEMORY ? s_cam : Cparam;  // one address space per instantiation
    const unsigned tile = s_tile;
    const int npx = A.w * A.h;
    const int i = (int)tile * TILE + threadIdx.x;
    Instance v{quiet_nan(), quiet_nan(), quiet_nan(), quiet_nan(), quiet_nan()};
    if (i < npx) v = pixel_instance(A, C, P, i);
    const bool valid = instance_valid(v);
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_counts[wid] = __popc(m);
    __syncthreads();
    if (wid == 0) {
        int cnt = 0;
        for (int k = 0; k < TILE / 32; k++) cnt += warp_counts[k];
        volatile unsigned long long* desc = S.desc;
        if (lane == 0)
            desc[tile] = scan_pack(S.epoch, tile == 0 ? SCAN_INCLUSIVE : SCAN_AGGREGATE, (unsigned)cnt);
        // look back over the predecessors, 32 at a time
        int prefix = 0;
        int j = (int)tile - 1 - lane;
        bool done = tile == 0;
        while (!done) {
            unsigned long long d = 0;
            bool ready;
            do {
                d = j >= 0 ? desc[j] : scan_pack(S.epoch, SCAN_INCLUSIVE, 0);
                ready = (unsigned)(d >> 34) == S.epoch && ((d >> 32) & 3u) != 0;
            } while (__any_sync(0xffffffffu, !ready));
            const bool incl = ((d >> 32) & 3u) == SCAN_INCLUSIVE;
            const unsigned incl_mask = __ballot_sync(0xffffffffu, incl);
            // lanes up to (and including) the nearest inclusive predecessor contribute
            const int first_incl = incl_mask ? __ffs(incl_mask) - 1 : 32;
            int contrib = lane <= first_incl ? (int)(unsigned)(d & 0xffffffffu) : 0;
            for (int o = 16; o >= 1; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            prefix += contrib;
            done = incl_mask != 0;
            j -= 32;
        }
        if (lane == 0) {
            if (tile != 0) desc[tile] = scan_pack(S.epoch, SCAN_INCLUSIVE, (unsigned)(prefix + cnt));
            s_prefix = prefix;
            if ((int)tile == (npx + TILE - 1) / TILE - 1) *total = prefix + cnt;
        }
    }
    __syncthreads();
    int rank = s_prefix + __popc(m & ((1u << lane) - 1u));
    for (int k = 0; k < wid; k++) rank += warp_counts[k];
    if (valid) {
        p2c[2 * (size_t)rank] = v.p2x, p2c[2 * (size_t)rank + 1] = v.p2y;
        p3c[3 * (size_t)rank] = v.p3x, p3c[3 * (size_t)rank + 1] = v.p3y, p3c[3 * (size_t)rank + 2] = v.p3z;
    }
}

// the sampler's constant uniform draws (reference: solve_batch_lambdatwist.cu:16-19,44-48)
__global__ void k_hypothesis_draws(float4* u4, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    curandStateXORWOW_t st;
    curand_init(233ULL, (unsigned long long)idx, 0, &st);
    float4 u;
    u.x = curand_uniform(&st);
    u.y = curand_uniform(&st);
    u.z = curand_uniform(&st);
    u.w = curand_uniform(&st);
    u4[idx] = u;
}

// ------------------------------------------------------------------------------------------------
// Quad-lane sampler: four consecutive lanes share one hypothesis (p3p_twist_quad.cuh).  Lane q of a quad fetches the
// q-th sampled correspondence (one gather per point instead of every thread gathering all four), the quad exchanges
// them with shuffles, every lane solves ITS candidate (plane, root) of the minimal problem, and the winner of the
// 4th-point test — found with four shuffles in the reference's scan order — converts its rotation and writes the
// hypothesis.  Same draws, same candidates, same arithmetic per candidate as one thread per hypothesis
// (reference: solve_batch_lambdatwist.cu:11-42), a quarter of the dependent chain.
// ------------------------------------------------------------------------------------------------
constexpr int kQuadBlock = 64;  // 16 hypotheses per block
template <int SOLVER>
__global__ void __launch_bounds__(kQuadBlock)
    k_solve_p3p_quad(const float* __restrict__ p2s, const float* __restrict__ p3s, const int* d_n_pts, int n_pts_host,
                     const float4* __restrict__ u4, float fx, float fy, float cx, float cy, float* rvecs, float* tvecs,
                     int n_poses) {
    const int gid = blockIdx.x * kQuadBlock + threadIdx.x;
    const int hyp = gid >> 2, slot = gid & 3;
    if (hyp >= n_poses) return;  // whole quads leave together (n_poses*4 threads are launched, rounded up to blocks)
    const int lane = threadIdx.x & 31, qbase = lane & ~3;
    const unsigned qmask = 0xFu << qbase;
    const int n_pts = d_n_pts ? *d_n_pts : n_pts_host;
    int best = -1;
    quad::Pose P;
    if (n_pts >= 4 || !d_n_pts) {
        const float4 u = u4[hyp];
        const float mine = slot == 0 ? u.x : slot == 1 ? u.y : slot == 2 ? u.z : u.w;
        const int i = (int)f_mul(mine, (float)n_pts);  // can equal n_pts when u == 1; buffers are padded
        const float pu = p2s[i * 2], pv = p2s[i * 2 + 1];
        const float px = p3s[i * 3], py = p3s[i * 3 + 1], pz = p3s[i * 3 + 2];
        float uv[8];
        quad::Vec3f X[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uv[2 * k] = __shfl_sync(qmask, pu, qbase + k);
            uv[2 * k + 1] = __shfl_sync(qmask, pv, qbase + k);
            X[k].x = __shfl_sync(qmask, px, qbase + k);
            X[k].y = __shfl_sync(qmask, py, qbase + k);
            X[k].z = __shfl_sync(qmask, pz, qbase + k);
        }
        float err = 0.f;
        const bool mine_exists = SOLVER == 0 ? quad::twist_lane(slot, uv, X, fx, fy, cx, cy, P, err)
                                             : quad::ap3p_lane(slot, uv, X, fx, fy, cx, cy, P, err);
        bool exists[4];
        float errs[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            exists[k] = __shfl_sync(qmask, (int)mine_exists, qbase + k) != 0;
            errs[k] = __shfl_sync(qmask, err, qbase + k);
        }
        best = quad::pick_by_fourth_point(exists, errs);
    }
    if (best < 0) {
        if (slot == 0) {
            const float nan = quiet_nan();
            rvecs[hyp * 3 + 0] = nan, rvecs[hyp * 3 + 1] = nan, rvecs[hyp * 3 + 2] = nan;
            tvecs[hyp * 3 + 0] = nan, tvecs[hyp * 3 + 1] = nan, tvecs[hyp * 3 + 2] = nan;
        }
        return;
    }
    if (slot != best) return;
    tvecs[hyp * 3 + 0] = P.t[0], tvecs[hyp * 3 + 1] = P.t[1], tvecs[hyp * 3 + 2] = P.t[2];
    float R[3][3], rv[3];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k / 3][k % 3] = P.R[k];
    rot::rotation_to_rvec(R, rv);
    rvecs[hyp * 3 + 0] = rv[0], rvecs[hyp * 3 + 1] = rv[1], rvecs[hyp * 3 + 2] = rv[2];
}

// order-preserving finite filter of the hypotheses (reference: voldor/geometry.cpp:156-165) fused with the
// rvec pre-scaling for mean-shift (geometry.cpp:191)
__global__ void __launch_bounds__(1024)
    k_filter_pool(const float* __restrict__ rvecs, const float* __restrict__ tvecs, int n_poses, float rvec_scale,
                  float* pool, int* used) {
    // each warp owns a contiguous slice; one barrier: slice counts -> slice offsets -> ordered writes
    __shared__ int warp_counts[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int per_warp = ((n_poses + 31) / 32 + 31) / 32 * 32;
    const int begin = wid * per_warp, end = min(n_poses, begin + per_warp);
    auto is_valid = [&](int i) {
        if (i >= end) return false;
        const float r0 = rvecs[i * 3], r1 = rvecs[i * 3 + 1], r2 = rvecs[i * 3 + 2];
        const float t0 = tvecs[i * 3], t1 = tvecs[i * 3 + 1], t2 = tvecs[i * 3 + 2];
        return (bool)isfinite(f_add(f_add(f_add(f_add(f_add(r0, r1), r2), t0), t1), t2));
    };
    int count = 0;
    for (int base = begin; base < end; base += 32) count += __popc(__ballot_sync(0xffffffffu, is_valid(base + lane)));
    if (lane == 0) warp_counts[wid] = count;
    __syncthreads();
    int offset = 0, total = 0;
    for (int k = 0; k < 32; k++) {
        if (k < wid) offset += warp_counts[k];
        total += warp_counts[k];
    }
    for (int base = begin; base < end; base += 32) {
        const int i = base + lane;
        const bool valid = is_valid(i);
        const unsigned m = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            float* o = pool + (size_t)(offset + __popc(m & ((1u << lane) - 1u))) * 6;
            o[0] = f_mul(rvecs[i * 3], rvec_scale), o[1] = f_mul(rvecs[i * 3 + 1], rvec_scale);
            o[2] = f_mul(rvecs[i * 3 + 2], rvec_scale);
            o[3] = tvecs[i * 3], o[4] = tvecs[i * 3 + 1], o[5] = tvecs[i * 3 + 2];
        }
        offset += __popc(m);
    }
    if (threadIdx.x == 0) *used = total;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int Collector::ensure(int w_, int h_, int N) {
    if (!stream) VB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    w = w_, h = h_;
    const int npx = w * h;
    if (npx > map_capacity) {
        if (p2_map) cudaFree(p2_map), cudaFree(p3_map), cudaFree(p2c), cudaFree(p3c);
        if (!d_count) VB_CUDA(cudaMalloc((void**)&d_count, sizeof(int)));
        VB_CUDA(cudaMalloc((void**)&p2_map, (size_t)npx * 2 * sizeof(float)));
        VB_CUDA(cudaMalloc((void**)&p3_map, (size_t)npx * 3 * sizeof(float)));
        VB_CUDA(cudaMalloc((void**)&p2c, ((size_t)npx + 1) * 2 * sizeof(float)));
        VB_CUDA(cudaMalloc((void**)&p3c, ((size_t)npx + 1) * 3 * sizeof(float)));
        VB_CUDA(cudaMemset(p2c, 0, ((size_t)npx + 1) * 2 * sizeof(float)));
        VB_CUDA(cudaMemset(p3c, 0, ((size_t)npx + 1) * 3 * sizeof(float)));
        const int nb = VB_DIV_CEIL(npx, kBlock);
        if (scan_desc) cudaFree(scan_desc);
        VB_CUDA(cudaMalloc((void**)&scan_desc, nb * sizeof(unsigned long long)));
        VB_CUDA(cudaMemset(scan_desc, 0, nb * sizeof(unsigned long long)));
        if (!scan_ticket) {
            VB_CUDA(cudaMalloc((void**)&scan_ticket, sizeof(unsigned int)));
            VB_CUDA(cudaMemset(scan_ticket, 0, sizeof(unsigned int)));
        }
        VB_CUDA(cudaDeviceSynchronize());  // the memsets run on the legacy stream, the kernels on `stream`
        map_capacity = npx;
    }
    (void)N;
    return 0;
}

void Collector::use_own_views() {
    flows = &flows_own;
    rig = rig_own.ptr, rig_pitch = rig_own.pitch, rig_plane = rig_own.layer_elems();
    depth = depth_own.ptr, depth_pitch = depth_own.pitch;
}

int Collector::collect(int N, const CollectParams& P, bool compact) {
    CollectView A;
    A.N = N, A.w = w, A.h = h;
    A.flows_tex = flows->tex;
    A.rig = rig, A.rig_pitch = rig_pitch, A.rig_plane = rig_plane;
    A.depth = depth, A.depth_pitch = depth_pitch;
    A.p2_map = p2_map, A.p3_map = p3_map;
    const int npx = w * h;
    const int nb = VB_DIV_CEIL(npx, kBlock);
    if (compact) {
        ScanState S;
        S.desc = scan_desc, S.ticket = scan_ticket;
        S.ticket_base = ticket_total, S.epoch = ++scan_epoch;
        if (scan_epoch >= (1u << 30)) scan_epoch = 0;  // descriptors hold 30 epoch bits; 0 is never issued
        const int nt = VB_DIV_CEIL(npx, kCompactTile);
        ticket_total += (unsigned)nt;
        if (d_cam)
            k_collect_compact<kCompactTile, true><<<nt, kCompactTile, 0, stream>>>(A, cam, d_cam, P, S, p2c, p3c, d_count);
        else
            k_collect_compact<kCompactTile, false><<<nt, kCompactTile, 0, stream>>>(A, cam, d_cam, P, S, p2c, p3c, d_count);
    } else {
        k_collect<<<nb, kBlock, 0, stream>>>(A, cam, P);
    }
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

const float4* HypothesisDraws::ensure(int n_poses, cudaStream_t s) {
    static std::mutex m;
    static float4* table = nullptr;  // superseded (smaller) tables stay allocated: kernels in flight may read them
    static int capacity = 0;
    std::lock_guard<std::mutex> lock(m);
    if (n_poses <= capacity) return table;
    const int cap = std::max(n_poses, 16384);
    float4* fresh = nullptr;
    if (cudaMalloc((void**)&fresh, (size_t)cap * sizeof(float4)) != cudaSuccess) return nullptr;
    k_hypothesis_draws<<<VB_DIV_CEIL(cap, 128), 128, 0, s>>>(fresh, cap);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess) return nullptr;
    table = fresh, capacity = cap;
    return table;
}

HypothesisDraws& global_draws() {
    static HypothesisDraws inst;
    return inst;
}

int solve_batch_p3p_device(const float* d_p3s, const float* d_p2s, const int* d_n_pts, int n_pts_host, float fx,
                           float fy, float cx, float cy, float* d_rvecs, float* d_tvecs, int n_poses,
                           bool use_ap3p, cudaStream_t s) {
    const float4* u4 = global_draws().ensure(n_poses, s);
    if (!u4) return (int)cudaErrorMemoryAllocation;
    const int nb = VB_DIV_CEIL(n_poses * 4, kQuadBlock);
    if (use_ap3p)
        k_solve_p3p_quad<1><<<nb, kQuadBlock, 0, s>>>(d_p2s, d_p3s, d_n_pts, n_pts_host, u4, fx, fy, cx, cy, d_rvecs,
                                                     d_tvecs, n_poses);
    else
        k_solve_p3p_quad<0><<<nb, kQuadBlock, 0, s>>>(d_p2s, d_p3s, d_n_pts, n_pts_host, u4, fx, fy, cx, cy, d_rvecs,
                                                     d_tvecs, n_poses);
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

int filter_pose_pool(const float* d_rvecs, const float* d_tvecs, int n_poses, float rvec_scale, float* d_pool,
                     int* d_used, cudaStream_t s) {
    k_filter_pool<<<1, 1024, 0, s>>>(d_rvecs, d_tvecs, n_poses, rvec_scale, d_pool, d_used);
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

}  // namespace vb
