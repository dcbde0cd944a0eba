// P3P instance collection, on-device compaction and batched pose hypotheses for sm_100a.
// See pose_sampler.cuh for the reference call sites this replaces.
#include "pose_sampler.cuh"
#include "geometry.cuh"
#include "p3p_lambdatwist.cuh"
#include "p3p_ap3p.cuh"
#include "rotation.cuh"
#include <curand_kernel.h>

namespace vb {

namespace {

constexpr int kBlock = 256;

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
    int* block_counts;
};

__device__ __forceinline__ float quiet_nan() { return __int_as_float(0x7fffffff); }  // CUDART_NAN_F

// one thread per pixel in raster order (reference: collect_p3p_instances.cu:70-145)
__global__ void __launch_bounds__(kBlock)
    k_collect(const CollectView A, const __grid_constant__ CamBlock C, const CollectParams P) {
    const int i = blockIdx.x * kBlock + threadIdx.x;
    const int npx = A.w * A.h;
    float p2x = quiet_nan(), p2y = quiet_nan();
    float p3x = quiet_nan(), p3y = quiet_nan(), p3z = quiet_nan();

    if (i < npx) {
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
        A.p2_map[2 * (size_t)i] = p2x;
        A.p2_map[2 * (size_t)i + 1] = p2y;
        A.p3_map[3 * (size_t)i] = p3x;
        A.p3_map[3 * (size_t)i + 1] = p3y;
        A.p3_map[3 * (size_t)i + 2] = p3z;
    }
    if (A.block_counts) {
        // validity rule of the host compaction loop (reference: voldor/geometry.cpp:72)
        const float s = f_add(f_add(f_add(f_add(p2x, p2y), p3x), p3y), p3z);
        const int cnt = __syncthreads_count(isfinite(s) ? 1 : 0);
        if (threadIdx.x == 0) A.block_counts[blockIdx.x] = cnt;
    }
}

// exclusive scan of per-block counts (single block)
__global__ void __launch_bounds__(1024) k_scan_counts(const int* counts, int* offsets, int nblocks, int* total) {
    __shared__ int warp_sums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < nblocks; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nblocks ? counts[i] : 0;
        int incl = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int ws = warp_sums[lane];
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += n;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const int prefix = carry + (wid > 0 ? warp_sums[wid - 1] : 0) + incl - v;
        if (i < nblocks) offsets[i] = prefix;
        __syncthreads();
        if (threadIdx.x == 1023) carry = prefix + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// scatter valid instances to their raster-order rank (reference: voldor/geometry.cpp:70-80)
__global__ void __launch_bounds__(kBlock)
    k_compact(const float* p2_map, const float* p3_map, const int* offsets, int npx, float* p2c, float* p3c) {
    __shared__ int warp_counts[kBlock / 32];
    const int i = blockIdx.x * kBlock + threadIdx.x;
    float a = quiet_nan(), b = 0, c = 0, d = 0, e = 0;
    if (i < npx) {
        a = p2_map[2 * (size_t)i], b = p2_map[2 * (size_t)i + 1];
        c = p3_map[3 * (size_t)i], d = p3_map[3 * (size_t)i + 1], e = p3_map[3 * (size_t)i + 2];
    }
    const bool valid = isfinite(f_add(f_add(f_add(f_add(a, b), c), d), e));
    const unsigned m = __ballot_sync(0xffffffffu, valid);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_counts[wid] = __popc(m);
    __syncthreads();
    int rank = offsets[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
    for (int k = 0; k < wid; k++) rank += warp_counts[k];
    if (valid) {
        p2c[2 * (size_t)rank] = a, p2c[2 * (size_t)rank + 1] = b;
        p3c[3 * (size_t)rank] = c, p3c[3 * (size_t)rank + 1] = d, p3c[3 * (size_t)rank + 2] = e;
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

// one thread = one hypothesis (reference: solve_batch_lambdatwist.cu:11-42, solve_batch_ap3p.cu:331-378)
template <bool AP3P>
__global__ void __launch_bounds__(32)
    k_solve_p3p(const float* __restrict__ p2s, const float* __restrict__ p3s, const int* d_n_pts, int n_pts_host,
                const float4* __restrict__ u4, float fx, float fy, float cx, float cy, float* rvecs, float* tvecs,
                int n_poses) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_poses) return;
    const int n_pts = d_n_pts ? *d_n_pts : n_pts_host;
    bool success = false;
    float R[3][3], t[3];
    if (n_pts >= 4 || !d_n_pts) {
        const float4 u = u4[idx];
        const float fn = (float)n_pts;
        const int i1 = (int)f_mul(u.x, fn);  // can equal n_pts when u == 1; buffers are padded
        const int i2 = (int)f_mul(u.y, fn);
        const int i3 = (int)f_mul(u.z, fn);
        const int i4 = (int)f_mul(u.w, fn);
        if (AP3P)
            success = ap3p::p4p_solve(&p2s[i1 * 2], &p2s[i2 * 2], &p2s[i3 * 2], &p2s[i4 * 2], &p3s[i1 * 3],
                                      &p3s[i2 * 3], &p3s[i3 * 3], &p3s[i4 * 3], fx, fy, cx, cy, R, t);
        else
            success = p3p::p4p_solve(&p2s[i1 * 2], &p2s[i2 * 2], &p2s[i3 * 2], &p2s[i4 * 2], &p3s[i1 * 3],
                                     &p3s[i2 * 3], &p3s[i3 * 3], &p3s[i4 * 3], fx, fy, cx, cy, R, t);
    }
    if (!success) {
        const float nan = quiet_nan();
        rvecs[idx * 3 + 0] = nan, rvecs[idx * 3 + 1] = nan, rvecs[idx * 3 + 2] = nan;
        tvecs[idx * 3 + 0] = nan, tvecs[idx * 3 + 1] = nan, tvecs[idx * 3 + 2] = nan;
        return;
    }
    tvecs[idx * 3 + 0] = t[0];
    tvecs[idx * 3 + 1] = t[1];
    tvecs[idx * 3 + 2] = t[2];
    float rv[3];
    rot::rotation_to_rvec(R, rv);
    rvecs[idx * 3 + 0] = rv[0];
    rvecs[idx * 3 + 1] = rv[1];
    rvecs[idx * 3 + 2] = rv[2];
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
        if (p2_map) cudaFree(p2_map), cudaFree(p3_map), cudaFree(p2c), cudaFree(p3c), cudaFree(block_counts),
            cudaFree(block_offsets);
        if (!d_count) VB_CUDA(cudaMalloc((void**)&d_count, sizeof(int)));
        VB_CUDA(cudaMalloc((void**)&p2_map, (size_t)npx * 2 * sizeof(float)));
        VB_CUDA(cudaMalloc((void**)&p3_map, (size_t)npx * 3 * sizeof(float)));
        VB_CUDA(cudaMalloc((void**)&p2c, ((size_t)npx + 1) * 2 * sizeof(float)));
        VB_CUDA(cudaMalloc((void**)&p3c, ((size_t)npx + 1) * 3 * sizeof(float)));
        VB_CUDA(cudaMemset(p2c, 0, ((size_t)npx + 1) * 2 * sizeof(float)));
        VB_CUDA(cudaMemset(p3c, 0, ((size_t)npx + 1) * 3 * sizeof(float)));
        const int nb = VB_DIV_CEIL(npx, kBlock);
        VB_CUDA(cudaMalloc((void**)&block_counts, nb * sizeof(int)));
        VB_CUDA(cudaMalloc((void**)&block_offsets, nb * sizeof(int)));
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
    A.block_counts = compact ? block_counts : nullptr;
    const int npx = w * h;
    const int nb = VB_DIV_CEIL(npx, kBlock);
    k_collect<<<nb, kBlock, 0, stream>>>(A, cam, P);
    VB_RETURN_IF_CUDA_ERROR();
    if (compact) {
        k_scan_counts<<<1, 1024, 0, stream>>>(block_counts, block_offsets, nb, d_count);
        k_compact<<<nb, kBlock, 0, stream>>>(p2_map, p3_map, block_offsets, npx, p2c, p3c);
        VB_RETURN_IF_CUDA_ERROR();
    }
    return 0;
}

Collector& global_collector() {
    static Collector inst;
    return inst;
}

int HypothesisDraws::ensure(int n_poses, cudaStream_t s) {
    if (n_poses <= capacity) return 0;
    if (u4) cudaFree(u4);
    u4 = nullptr, capacity = 0;
    VB_CUDA(cudaMalloc((void**)&u4, (size_t)n_poses * sizeof(float4)));
    k_hypothesis_draws<<<VB_DIV_CEIL(n_poses, 128), 128, 0, s>>>(u4, n_poses);
    VB_RETURN_IF_CUDA_ERROR();
    VB_CUDA(cudaStreamSynchronize(s));  // other streams may consume the table
    capacity = n_poses;
    return 0;
}

HypothesisDraws& global_draws() {
    static HypothesisDraws inst;
    return inst;
}

int solve_batch_p3p_device(const float* d_p3s, const float* d_p2s, const int* d_n_pts, int n_pts_host, float fx,
                           float fy, float cx, float cy, float* d_rvecs, float* d_tvecs, int n_poses,
                           bool use_ap3p, cudaStream_t s) {
    if (int e = global_draws().ensure(n_poses, s)) return e;
    const int nb = VB_DIV_CEIL(n_poses, 32);
    if (use_ap3p)
        k_solve_p3p<true><<<nb, 32, 0, s>>>(d_p2s, d_p3s, d_n_pts, n_pts_host, global_draws().u4, fx, fy, cx, cy,
                                            d_rvecs, d_tvecs, n_poses);
    else
        k_solve_p3p<false><<<nb, 32, 0, s>>>(d_p2s, d_p3s, d_n_pts, n_pts_host, global_draws().u4, fx, fy, cx, cy,
                                             d_rvecs, d_tvecs, n_poses);
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
