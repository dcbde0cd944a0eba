// Mean-shift and robust Gaussian fit kernels for sm_100a.  See pose_mode.cuh.
//
// Both run as ONE launch of ONE thread-block cluster (8 CTAs x 512 threads, distributed over 8 SMs): the pose
// pool (<= 8192 x 6 floats) is split by 512-element tree blocks over the CTAs and staged in their shared memory;
// every iteration of the fixed-point loop runs on the device — weights, the reference-ordered tree sums
// (tree_sum.cuh; every CTA stores its level-1 partials straight into the partial table of all 8 CTAs through
// distributed shared memory, one hardware cluster barrier per iteration, double buffered; pools beyond 32768
// elements exchange through L2 instead), the host-side arithmetic of the reference's loop (mean update /
// displacement test; 6x6 FP64 LU, inverse and shrinkage for the robust fit — evaluated redundantly and
// identically by every CTA) and the convergence test.  One launch replaces the reference's up to 100 x (kernel + 2..3 multi-pass reductions +
// 2..4 blocking copies) per call.
#include "pose_mode.cuh"
#include "depth_em.cuh"
#include "residual_model.cuh"
#include "tree_sum.cuh"
#include "libc_rand.h"
#include "host_math.h"
#include <cmath>
#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace vb {

namespace {

constexpr int kCluster = 8;
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxTreeBlocks = 512;  // pool size limit 512*512 (two tree levels)
constexpr int kDsmemBlocks = 64;     // pools up to 64*512 elements exchange partial sums through DSMEM
constexpr int kMaxTrialBatch = 32;   // start-sample trials evaluated inside one launch
constexpr int kMeanshiftQ = kMaxTrialBatch > kMeanshiftMaxDims + 1 ? kMaxTrialBatch : kMeanshiftMaxDims + 1;
constexpr size_t kSmemBudget = 200 * 1024;

struct PoolSource {
    // either a ready pool (space + n from device or host) or raw hypotheses to filter first
    const float* space;
    const int* d_n;
    int n_host;
    const float* rvecs;  // optional fused finite-filter (reference voldor/geometry.cpp:156-165,191)
    const float* tvecs;
    int n_poses;
    float rvec_scale;
    float* pool_out;
    int* used_out;
    int* cta_counts;       // [kCluster] scratch
    const int* aux_count;  // optional device counter echoed into the result (saves the caller a separate copy)
};

struct MeanshiftArgs {
    float io_mean[kMeanshiftMaxDims];  // caller's mean: first displacement is measured against it (Q13)
    int center_idx;                    // >= 0: start from space[center_idx]; < 0: start from io_mean
    int trial_only;                    // 1: only the kernel-weight sum around the start point
    int dims;
    float kernel_var, epsilon;
    int max_iters;
    int slice_in_smem;
    // start-sample selection fused into the launch (reference host loop meanshift.cu:73-97): n_trials > 0
    int n_trials;
    int trial_idx[kMaxTrialBatch];
    float good_init_confidence;
    PoolSource src;
    PoseTail tail;
    long long* phase_cycles;  // optional [16]: clock64 totals of rank 0 at [8..15] (profiling)
};

// camera pose from the mean-shift mode (reference voldor/geometry.cpp:247-262 for a successive pose without robust
// refinement); mirrors Window::apply_camera_result on the host
__device__ void pose_tail(const PoseTail& T, const float* mode, int dims, int n_points, int pool_used,
                          MeanshiftResult* out) {
    int ok = (n_points >= 4 && pool_used > 0 && dims == 6) ? 1 : 0;
    float pose[6] = {0, 0, 0, 0, 0, 0};
    if (ok) {
        for (int d = 0; d < 6; d++) pose[d] = mode[d];
        for (int d = 0; d < 3; d++) pose[d] = f_mul(pose[d], T.inv_rvec_scale);
        for (int d = 0; d < 6; d++)
            if (!isfinite(pose[d])) ok = 0;  // cv::checkRange
    }
    out->ok = ok;
    if (!ok) return;
    float R[9];
    hm::rvec_to_matrix(pose, R);
    for (int k = 0; k < 9; k++) out->R[k] = R[k], T.d_cams->R[T.cam_index][k] = R[k];
    for (int k = 0; k < 3; k++) out->t[k] = pose[3 + k], T.d_cams->t[T.cam_index][k] = pose[3 + k];
}

// Ordered compaction of finite hypotheses across the whole cluster; returns the pool size on every thread.
__device__ int build_pool(const PoolSource& S, cg::cluster_group& cluster, int rank) {
    __shared__ int s_scan[kWarps];
    __shared__ int s_base, s_total;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gid = rank * kThreads + threadIdx.x;
    const int E = (S.n_poses + kCluster * kThreads - 1) / (kCluster * kThreads);
    const int lo = min(S.n_poses, gid * E), hi = min(S.n_poses, lo + E);
    auto valid = [&](int i) {
        const float* r = S.rvecs + (size_t)i * 3;
        const float* t = S.tvecs + (size_t)i * 3;
        return (bool)isfinite(f_add(f_add(f_add(f_add(f_add(r[0], r[1]), r[2]), t[0]), t[1]), t[2]));
    };
    int cnt = 0;
    for (int i = lo; i < hi; i++) cnt += valid(i) ? 1 : 0;
    int incl = cnt;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int k = 0; k < warp; k++) base += s_scan[k];
    if (threadIdx.x == kThreads - 1) {
        S.cta_counts[rank] = base + incl;
        __threadfence();
    }
    cluster.sync();
    if (threadIdx.x == 0) {
        int b = 0, tot = 0;
        for (int k = 0; k < kCluster; k++) {
            const int c = ((volatile int*)S.cta_counts)[k];
            if (k < rank) b += c;
            tot += c;
        }
        s_base = b, s_total = tot;
    }
    __syncthreads();
    int pos = s_base + base + incl - cnt;
    for (int i = lo; i < hi; i++) {
        if (valid(i)) {
            const float* r = S.rvecs + (size_t)i * 3;
            const float* t = S.tvecs + (size_t)i * 3;
            float* o = S.pool_out + (size_t)pos * 6;
            o[0] = f_mul(r[0], S.rvec_scale), o[1] = f_mul(r[1], S.rvec_scale), o[2] = f_mul(r[2], S.rvec_scale);
            o[3] = t[0], o[4] = t[1], o[5] = t[2];
            pos++;
        }
    }
    __threadfence();
    cluster.sync();
    if (rank == 0 && threadIdx.x == 0) *S.used_out = s_total;
    return s_total;
}

// Same compaction without any exchange: every CTA scans the validity of ALL hypotheses itself (the table is small
// and L2 resident: 8192 x 24 B), so each one knows every output position, and then loads only the hypotheses whose
// tree block it owns — straight into its own shared-memory slice and into its part of the global pool (kept for later
// consumers such as the robust fit).  Measured against exchanging positions/elements through distributed shared
// memory (scalar remote stores: 16k cycles per launch) this takes ~3k.  At most 32 hypotheses per thread.
// `smem` is this CTA's dynamic shared memory (layout: weights [n_local], slice [n_local][6], ...).
__device__ int build_pool_local(const PoolSource& S, int rank, float* smem) {
    // warp w scans the contiguous chunk [w*C, w*C + C) of hypotheses, lane l those at chunk + 32 j + l: coalesced,
    // independent loads (memory-level parallelism), validity bit j kept in a per-lane mask
    __shared__ int s_scan[kWarps];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int J = (S.n_poses + kThreads - 1) / kThreads;  // rounds per lane, <= 32
    const int chunk = warp * J * 32;
    unsigned mask = 0;
#pragma unroll 8
    for (int j = 0; j < J; j++) {
        const int i = chunk + j * 32 + lane;
        if (i < S.n_poses) {
            const float* r = S.rvecs + (size_t)i * 3;
            const float* t = S.tvecs + (size_t)i * 3;
            const float r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
            const float t0 = __ldg(t), t1 = __ldg(t + 1), t2 = __ldg(t + 2);
            if (isfinite(f_add(f_add(f_add(f_add(f_add(r0, r1), r2), t0), t1), t2))) mask |= 1u << j;
        }
    }
    int warp_total = __popc(mask);
    for (int o = 16; o >= 1; o >>= 1) warp_total += __shfl_xor_sync(0xffffffffu, warp_total, o);
    if (lane == 0) s_scan[warp] = warp_total;
    __syncthreads();
    int base = 0, total = 0;
    for (int k = 0; k < kWarps; k++) {
        const int c = s_scan[k];
        if (k < warp) base += c;
        total += c;
    }
    const int NB = (total + 511) / 512;
    const int my_nlb = max(0, (NB - rank + kCluster - 1) / kCluster);
    float* slice = smem + (size_t)my_nlb * 512;
    int running = base;
    // second pass in groups of 4 rounds: the (L2-resident) reloads of a whole group are issued before any of its
    // stores, otherwise every round would pay a full load latency
    for (int j0 = 0; j0 < J; j0 += 4) {
        float x[4][6];
        bool valid[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int j = j0 + u;
            valid[u] = j < J && ((mask >> j) & 1u);
            if (valid[u]) {
                const int i = chunk + j * 32 + lane;
                const float* r = S.rvecs + (size_t)i * 3;
                const float* t = S.tvecs + (size_t)i * 3;
                x[u][0] = __ldg(r), x[u][1] = __ldg(r + 1), x[u][2] = __ldg(r + 2);
                x[u][3] = __ldg(t), x[u][4] = __ldg(t + 1), x[u][5] = __ldg(t + 2);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const unsigned bal = __ballot_sync(0xffffffffu, valid[u]);
            const int pos = running + __popc(bal & ((1u << lane) - 1u));
            running += __popc(bal);
            if (valid[u] && ((pos >> 9) % kCluster) == rank) {
                float* g = S.pool_out + (size_t)pos * 6;
                float* dst = slice + ((size_t)((pos >> 9) / kCluster) * 512 + (pos & 511)) * 6;
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    const float v = d < 3 ? f_mul(x[u][d], S.rvec_scale) : x[u][d];
                    g[d] = v, dst[d] = v;
                }
            }
        }
    }
    __syncthreads();
    if (rank == 0 && threadIdx.x == 0) *S.used_out = total;
    return total;
}

// element bookkeeping of one CTA: it owns tree blocks b = rank, rank+8, ... ; local block lb = b / 8
struct Slice {
    int N, NB, nlb, rank, dims;
    const float* global;  // pool in global memory
    const float* local;   // staged slice (or nullptr)
    __device__ __forceinline__ int gidx(int li) const { return ((li >> 9) * kCluster + rank) * 512 + (li & 511); }
    __device__ __forceinline__ float x(int li, int d) const {
        return local ? local[(size_t)li * dims + d] : ((const volatile float*)global)[(size_t)gidx(li) * dims + d];
    }
};

__device__ void stage_slice(Slice& S, float* smem_pool, bool in_smem) {
    S.local = nullptr;
    if (!in_smem) return;
    const int n_local = S.nlb * 512;
    for (int k = threadIdx.x; k < n_local * S.dims; k += kThreads) {
        const int li = k / S.dims, d = k - li * S.dims;
        const int g = S.gidx(li);
        smem_pool[k] = (g < S.N) ? ((const volatile float*)S.global)[(size_t)g * S.dims + d] : 0.f;
    }
    S.local = smem_pool;
}

// Where the level-1 partial sums of one iteration live: a [Q][stride] table in the shared memory of every CTA
// (DSMEM exchange) or in global memory.
struct Exchange {
    float* table;
    int stride;
    bool dsmem;
};
__device__ __forceinline__ Exchange exchange_for(int iter, int NB, float* s_table, int q_max, float* partials_g) {
    Exchange X;
    X.dsmem = NB <= kDsmemBlocks;
    X.stride = X.dsmem ? kDsmemBlocks : kMaxTreeBlocks;
    X.table = (X.dsmem ? s_table : partials_g) + (size_t)(iter & 1) * q_max * X.stride;
    return X;
}
__device__ __forceinline__ void exchange_sync(const Exchange& X, cg::cluster_group& cluster) {
    if (!X.dsmem) __threadfence();
    cluster.sync();  // barrier.cluster arrive.release / wait.acquire: remote shared-memory stores are visible after it
}

// level 1: one warp per (local tree block, quantity); level 2: one warp per quantity (every CTA, redundantly)
template <class Val>
__device__ __forceinline__ void tree_level1(const Slice& S, int Q, const Exchange& X, cg::cluster_group& cluster,
                                            Val val) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int task = warp; task < S.nlb * Q; task += kWarps) {
        const int q = task / S.nlb, lb = task % S.nlb;
        const int b = lb * kCluster + S.rank;
        if (b >= S.NB) continue;
        const int count = min(512, S.N - b * 512);
        float v = tree_sum_512([&](int i) { return val(q, lb * 512 + i); }, count, lane);
        if (lane == 0 && S.N == 1) v = val(q, 0);  // the reference launches no reduction at all for a single element
        float* dst = X.table + q * X.stride + b;
        if (X.dsmem) {
            v = __shfl_sync(0xffffffffu, v, 0);
            if (lane < kCluster) *cluster.map_shared_rank(dst, lane) = v;
        } else if (lane == 0) {
            *dst = v;
        }
    }
}
__device__ __forceinline__ void tree_level2(int NB, int Q, const Exchange& X, float* sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = warp; q < Q; q += kWarps) {
        const volatile float* p = X.table + q * X.stride;
        const float v = (NB == 1) ? p[0] : tree_sum_512([&](int i) { return (float)p[i]; }, NB, lane);
        if (lane == 0) sums[q] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// mean-shift (reference: meanshift.cu:12-31 weights, :99-134 host loop)
// ------------------------------------------------------------------------------------------------
// FAST6: dims == 6 with the slice and the per-element products [Q][n_local] resident in shared memory — the
// element loop is fully unrolled and every tree sum reads one conflict-free row.
template <bool FAST6>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
    k_meanshift(const MeanshiftArgs A, float* partials_g, MeanshiftResult* out) {
    extern __shared__ float smem[];
    __shared__ float c_mean[kMeanshiftMaxDims];
    __shared__ float io_mean[kMeanshiftMaxDims];
    __shared__ float sums[kMeanshiftQ];
    __shared__ float s_part[2 * kMeanshiftQ * kDsmemBlocks];
    __shared__ float s_trial[kMaxTrialBatch * kMeanshiftMaxDims];
    __shared__ int s_center_idx, s_trials_used;
    __shared__ int done;
    __shared__ int s_used_iters;
    __shared__ float s_confidence, s_wsum;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dims = A.dims;
    const int Q = A.trial_only ? 1 : dims + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    long long tick = clock64(), ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const bool profiling = A.phase_cycles != nullptr && rank == 0 && threadIdx.x == 0;
#define VB_PHASE(k)                           \
    do {                                      \
        if (profiling) {                      \
            const long long now_ = clock64(); \
            ph[k] += now_ - tick;             \
            tick = now_;                      \
        }                                     \
    } while (0)
    int N;
    const float* space_g = A.src.space;
    bool staged = false;
    if (A.src.rvecs) {
        // dims == 6 here by construction (rvec, tvec)
        staged = A.slice_in_smem && A.src.n_poses <= 32 * kThreads;
        if (staged) {
            N = build_pool_local(A.src, rank, smem);
            cluster.sync();  // every CTA of the cluster is resident before anyone stores into its shared memory
        } else {
            N = build_pool(A.src, cluster, rank);
        }
        space_g = A.src.pool_out;
    } else {
        N = A.src.d_n ? *A.src.d_n : A.src.n_host;
        cluster.sync();  // every CTA of the cluster is resident before anyone stores into its shared memory
    }
    if (N <= 0) {
        if (rank == 0 && threadIdx.x == 0) {
            out->used_iters = 0, out->n = N, out->weight_sum = 0.f, out->confidence = 0.f;
            out->aux_count = A.src.aux_count ? *A.src.aux_count : 0;
            out->ok = 0;
        }
        return;
    }
    Slice S;
    S.N = N, S.NB = (N + 511) / 512, S.rank = rank, S.dims = dims, S.global = space_g;
    S.nlb = (S.NB - rank + kCluster - 1) / kCluster;
    if (S.nlb < 0) S.nlb = 0;
    float* wv = smem;  // weights of the local slice
    VB_PHASE(0);
    if (staged)
        S.local = smem + (size_t)S.nlb * 512;
    else
        stage_slice(S, smem + (size_t)S.nlb * 512, A.slice_in_smem != 0);
    const int n_local = S.nlb * 512;
    float* prod = smem + (size_t)n_local * (1 + dims);  // FAST6: rows w, w*x_0 .. w*x_5
    const float two_var = f_add(A.kernel_var, A.kernel_var);  // 2*kernel_var
    if (threadIdx.x == 0) s_center_idx = A.center_idx, s_trials_used = 0;
    if (A.n_trials > 0) {
        // weight sum around every trial sample, then the reference's selection loop
        const int T = A.n_trials;
        for (int k = threadIdx.x; k < T * dims; k += kThreads)
            s_trial[k] = ((const volatile float*)space_g)[(size_t)A.trial_idx[k / dims] * dims + (k % dims)];
        __syncthreads();
        const Exchange X = exchange_for(1, S.NB, s_part, kMeanshiftQ, partials_g);
        tree_level1(S, T, X, cluster, [&](int q, int li) {
            float l2 = 0.f;
            for (int d = 0; d < dims; d++) {
                const float diff = f_sub(S.x(li, d), s_trial[q * dims + d]);
                l2 = f_fma(diff, diff, l2);
            }
            return expf(f_div(-l2, two_var));
        });
        exchange_sync(X, cluster);
        tree_level2(S.NB, T, X, sums);
        __syncthreads();
        if (threadIdx.x == 0) {
            float best_conf = 0.f;
            int best_idx = -1, used = T;
            const float good = f_mul(A.good_init_confidence, (float)N);
            for (int trial = 0; trial < T; trial++) {
                if (sums[trial] > best_conf) best_conf = sums[trial], best_idx = A.trial_idx[trial];
                if (best_conf > good) {
                    used = trial + 1;
                    break;
                }
            }
            s_center_idx = best_idx < 0 ? 0 : best_idx;
            s_trials_used = used;
        }
    }
    __syncthreads();
    if (threadIdx.x < dims) {
        io_mean[threadIdx.x] = A.io_mean[threadIdx.x];
        c_mean[threadIdx.x] = (s_center_idx >= 0)
                                  ? ((const volatile float*)space_g)[(size_t)s_center_idx * dims + threadIdx.x]
                                  : A.io_mean[threadIdx.x];
    }
    if (threadIdx.x == 0) done = 0, s_used_iters = 0, s_confidence = 0.f, s_wsum = 0.f;
    __syncthreads();
    VB_PHASE(1);

    const int n_iters = A.trial_only ? 1 : A.max_iters;

    for (int iter = 0; iter < n_iters; iter++) {
        const Exchange X = exchange_for(iter, S.NB, s_part, kMeanshiftQ, partials_g);
        // weights w_i = exp(-|x_i - mu|^2 / (2 var)) of the local slice
        if constexpr (FAST6) {
            float cm[6];
#pragma unroll
            for (int d = 0; d < 6; d++) cm[d] = c_mean[d];
#pragma unroll 2
            for (int li = threadIdx.x; li < n_local; li += kThreads) {
                float wgt = 0.f, x[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                if (S.gidx(li) < N) {
                    float l2 = 0.f;
#pragma unroll
                    for (int d = 0; d < 6; d++) {
                        x[d] = S.local[(size_t)li * 6 + d];
                        const float diff = f_sub(x[d], cm[d]);
                        l2 = f_fma(diff, diff, l2);
                    }
                    wgt = expf(f_div(-l2, two_var));
                }
                prod[li] = wgt;
                if (Q > 1) {
#pragma unroll
                    for (int d = 0; d < 6; d++) prod[(size_t)(1 + d) * n_local + li] = f_mul(wgt, x[d]);
                }
            }
            __syncthreads();
            VB_PHASE(2);
            tree_level1(S, Q, X, cluster, [&](int q, int li) { return prod[(size_t)q * n_local + li]; });
        } else {
            for (int li = threadIdx.x; li < n_local; li += kThreads) {
                float wgt = 0.f;
                if (S.gidx(li) < N) {
                    float l2 = 0.f;
                    for (int d = 0; d < dims; d++) {
                        const float diff = f_sub(S.x(li, d), c_mean[d]);
                        l2 = f_fma(diff, diff, l2);
                    }
                    wgt = expf(f_div(-l2, two_var));
                }
                wv[li] = wgt;
            }
            __syncthreads();
            tree_level1(S, Q, X, cluster,
                        [&](int q, int li) { return q == 0 ? wv[li] : f_mul(wv[li], S.x(li, q - 1)); });
        }
        VB_PHASE(3);
        exchange_sync(X, cluster);
        VB_PHASE(4);
        tree_level2(S.NB, Q, X, sums);
        __syncthreads();
        VB_PHASE(5);
        if (warp == 0) {
            // host part of the reference iteration (meanshift.cu:112-133), one lane per dimension
            const float wsum = sums[0];
            if (A.trial_only) {
                if (lane == 0) s_wsum = wsum;
            } else {
                float mean_new = 0.f, sq = 0.f;
                if (lane < dims) {
                    mean_new = f_div(sums[lane + 1], wsum);
                    const float df = f_sub(io_mean[lane], mean_new);
                    sq = f_mul(df, df);
                }
                float disp = 0.f;
                for (int d = 0; d < dims; d++) disp = f_add(disp, __shfl_sync(0xffffffffu, sq, d));
                disp = __fsqrt_rn(disp);
                if (lane < dims) io_mean[lane] = mean_new, c_mean[lane] = mean_new;
                if (lane == 0) {
                    s_wsum = wsum;
                    s_confidence = f_div(wsum, (float)N);
                    s_used_iters = iter + 1;
                    if (disp < A.epsilon) done = 1;
                }
            }
        }
        __syncthreads();
        VB_PHASE(6);
        if (done) break;
    }
    if (rank == 0 && threadIdx.x == 0) {
        for (int d = 0; d < dims; d++) out->mean[d] = io_mean[d];
        out->confidence = s_confidence;
        out->weight_sum = s_wsum;
        out->used_iters = s_used_iters;
        out->trials_used = s_trials_used;
        out->aux_count = A.src.aux_count ? *A.src.aux_count : 0;
        out->n = N;
        if (A.tail.d_cams) pose_tail(A.tail, io_mean, dims, out->aux_count, N, out);
        VB_PHASE(7);
        if (A.phase_cycles)
            for (int k = 0; k < 8; k++) A.phase_cycles[8 + k] += ph[k];
    }
#undef VB_PHASE
}

// ------------------------------------------------------------------------------------------------
// robust Gaussian fit, whole EM loop on the device
// (reference: fit_robust_gaussian.cu:56-97 E-step, :160-263 host loop, aux_funs.cpp:101-141 6x6 FP64 helpers)
// ------------------------------------------------------------------------------------------------
struct RobustArgs {
    float mean[kRobustMaxDims];
    float covar[21];  // lower-triangular packed start covariance
    float trunc_sigma, scale, covar_reg_lambda, epsilon;
    int N, dims, max_iters;
    int slice_in_smem, centred_in_smem;
    const float* space;
    long long* phase_cycles;  // optional [8]: per-phase clock64 totals of rank 0 (debug / profiling)
};
struct RobustResult {
    float mean[kRobustMaxDims];
    float covar[21];
    float density;
    int used_iters;
    int reliable;
};

// FP64 helpers with every operation individually rounded (the host code they mirror is compiled without FMA)
__device__ __forceinline__ double d_mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double d_add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double d_sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double d_div(double a, double b) { return __ddiv_rn(a, b); }

// 6x6 covariance step of one robust-fit iteration on ONE warp, the augmented matrix [a | b] held in registers: lane
// c < 6 owns column c of a, lane 6 + c column c of b; pivot rows, multipliers and the upper triangle travel by
// shuffles.  Operation for operation the sequential host code it mirrors (aux shim / cv::Matx66d: LU with partial
// pivoting, det, back-substitution), every FP64 op individually rounded.  s_cov (packed lower triangle, float) is
// replaced by its regularised version, s_cinv receives the inverse; returns false when det <= 0.
__device__ __forceinline__ bool covariance_step_6x6(int lane, float* s_cov, float* s_cinv, bool shrink, double lambda) {
    const unsigned full = 0xffffffffu;
    const bool is_a = lane < 6, is_b = lane >= 6 && lane < 12;
    const int c = is_a ? lane : lane - 6;
    double col[6];
#pragma unroll
    for (int r = 0; r < 6; r++) {
        col[r] = 0.0;
        if (is_a) {
            const int hi = r > c ? r : c, lo = r > c ? c : r;
            col[r] = (double)s_cov[(hi * hi + hi) / 2 + lo];
        } else if (is_b) {
            col[r] = (r == c) ? 1.0 : 0.0;
        }
    }
    if (shrink) {  // towards tr/n * I (Ledoit-Wolf with a fixed lambda)
        double tr = 0.0;
#pragma unroll
        for (int i = 0; i < 6; i++) tr = d_add(tr, __shfl_sync(full, col[i], i));
        const double m = d_div(tr, 6.0);
        const double lm = d_mul(lambda, m), oml = d_sub(1.0, lambda);
        if (is_a) {
#pragma unroll
            for (int r = 0; r < 6; r++) col[r] = d_add(d_mul(lm, r == c ? 1.0 : 0.0), d_mul(oml, col[r]));
        }
    }
    __syncwarp();  // every lane has read s_cov before anybody overwrites it
    if (is_a) {  // the regularised covariance is the one reported on convergence (fit_robust_gaussian.cu:203)
#pragma unroll
        for (int r = 0; r < 6; r++)
            if (r >= c) s_cov[(r * r + r) / 2 + c] = (float)col[r];
    }
    double det = 1.0;
    bool singular = false;
#pragma unroll
    for (int i = 0; i < 6; i++) {
        if (singular) continue;
        // partial pivoting on column i (owned by lane i): first maximum, strict '>'
        int k = i;
        double best = fabs(col[i]);
#pragma unroll
        for (int j = i + 1; j < 6; j++) {
            const double v = fabs(col[j]);
            if (v > best) best = v, k = j;
        }
        k = __shfl_sync(full, k, i);
        best = __shfl_sync(full, best, i);
        if (best < 2.220446049250313e-16 * 100) {
            singular = true;
            continue;
        }
        if (k != i) {
            if ((is_a && c >= i) || is_b) {
#pragma unroll
                for (int j = i + 1; j < 6; j++)
                    if (k == j) {
                        const double tmp = col[i];
                        col[i] = col[j];
                        col[j] = tmp;
                    }
            }
            det = -det;
        }
        const double piv = __shfl_sync(full, col[i], i);
        const double d = d_div(-1.0, piv);
#pragma unroll
        for (int r = i + 1; r < 6; r++) {
            const double alpha = d_mul(__shfl_sync(full, col[r], i), d);
            if ((is_a && c > i) || is_b) col[r] = d_add(col[r], d_mul(alpha, col[i]));
        }
        det = d_mul(det, piv);
    }
    if (singular) det = 0.0;
    if (!(det > 0)) return false;  // inverse only written for det > 0 (aux_funs.cpp:104-111)
#pragma unroll
    for (int i = 5; i >= 0; i--) {
        const double uii = __shfl_sync(full, col[i], i);
        double acc = col[i];
#pragma unroll
        for (int k = i + 1; k < 6; k++) acc = d_sub(acc, d_mul(__shfl_sync(full, col[i], k), col[k]));
        if (is_b) col[i] = d_div(acc, uii);
    }
    if (is_b) {
#pragma unroll
        for (int r = 0; r < 6; r++)
            if (r >= c) s_cinv[(r * r + r) / 2 + c] = (float)col[r];
    }
    return true;
}

template <bool FAST6>
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
    k_robust_fit(const RobustArgs A, float* partials_g, RobustResult* out) {
    extern __shared__ float smem[];
    __shared__ float s_mean[kRobustMaxDims];
    __shared__ float s_cov[21], s_cinv[21];
    __shared__ float sums[28];
    __shared__ float s_part[2 * 28 * kDsmemBlocks];
    __shared__ int s_d1[28], s_d2[28];  // (row, column) of the packed lower-triangular covariance terms
    __shared__ int s_lu_state;          // 0 ok, 2 unreliable (det <= 0)
    // 6x6 FP64 scratch of the covariance step: LU with partial pivoting on the augmented matrix [a | b], run by
    // warp 0 with one lane per column — the same operations on the same elements as the sequential host code it
    // mirrors (aux shim / cv::Matx66d), only issued in parallel
    __shared__ double sa[36], sb[36];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int N = A.N, dims = A.dims;
    const int cdims = (dims * dims + dims) / 2;
    const int Q = 1 + dims + cdims;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;

    cluster.sync();  // every CTA of the cluster is resident before anyone stores into its shared memory
    Slice S;
    S.N = N, S.NB = (N + 511) / 512, S.rank = rank, S.dims = dims, S.global = A.space;
    S.nlb = (S.NB - rank + kCluster - 1) / kCluster;
    if (S.nlb < 0) S.nlb = 0;
    const int n_local = S.nlb * 512;
    float* wv = smem;
    stage_slice(S, smem + (size_t)n_local, A.slice_in_smem != 0);
    // centred, scaled coordinates of the local slice (rewritten by every E-step)
    float* av = (!FAST6 && A.centred_in_smem) ? smem + (size_t)n_local * (1 + (A.slice_in_smem ? dims : 0)) : nullptr;
    float* prod = smem + (size_t)n_local * (1 + dims);  // FAST6: rows w, w*x_d (6), w*a_d1*a_d2 (21)
    if (t == 0) {
        for (int d = 0; d < dims; d++) s_mean[d] = A.mean[d];
        for (int k = 0; k < cdims; k++) s_cov[k] = A.covar[k];
        for (int d1 = 0; d1 < dims; d1++)
            for (int d2 = 0; d2 <= d1; d2++) s_d1[(d1 * d1 + d1) / 2 + d2] = d1, s_d2[(d1 * d1 + d1) / 2 + d2] = d2;
    }
    __syncthreads();

    float weight = 0.f;  // sum of weights of the previous E-step (uniform over all threads)
    int state = 0;       // 0 continue, 1 converged, 2 unreliable (uniform)
    int iter = 0;
    long long tick = clock64(), ph[6] = {0, 0, 0, 0, 0, 0};
    const bool profiling = A.phase_cycles != nullptr && rank == 0 && t == 0;
#define VB_PHASE(k)                           \
    do {                                      \
        if (profiling) {                      \
            const long long now_ = clock64(); \
            ph[k] += now_ - tick;             \
            tick = now_;                      \
        }                                     \
    } while (0)
    for (iter = 0; iter < A.max_iters; iter++) {
        const Exchange X = exchange_for(iter, S.NB, s_part, 28, partials_g);
        if (FAST6 && warp == 0) {
            const bool ok = covariance_step_6x6(lane, s_cov, s_cinv, iter > 0 && A.covar_reg_lambda > 0,
                                                (double)A.covar_reg_lambda);
            if (lane == 0) s_lu_state = ok ? 0 : 2;
        } else if (warp == 0) {
            // half -> full (double); shrink towards tr/n * I from the 2nd iteration on
            for (int e = lane; e < 36; e += 32) {
                const int r = e / 6, c = e % 6;
                if (r < dims && c < dims) {
                    const int hi = r > c ? r : c, lo = r > c ? c : r;
                    sa[e] = (double)s_cov[(hi * hi + hi) / 2 + lo];
                    sb[e] = (r == c) ? 1.0 : 0.0;
                }
            }
            __syncwarp();
            if (iter > 0 && A.covar_reg_lambda > 0) {
                const double lambda = (double)A.covar_reg_lambda;
                double tr = 0;
                for (int i = 0; i < dims; i++) tr = d_add(tr, sa[i * 6 + i]);
                const double m = d_div(tr, (double)dims);
                const double lm = d_mul(lambda, m), oml = d_sub(1.0, lambda);
                __syncwarp();
                for (int e = lane; e < 36; e += 32) {
                    const int r = e / 6, c = e % 6;
                    if (r < dims && c < dims) sa[e] = d_add(d_mul(lm, r == c ? 1.0 : 0.0), d_mul(oml, sa[e]));
                }
                __syncwarp();
            }
            // the regularised covariance is the one reported on convergence (fit_robust_gaussian.cu:203)
            for (int e = lane; e < 36; e += 32) {
                const int r = e / 6, c = e % 6;
                if (r < dims && c <= r) s_cov[(r * r + r) / 2 + c] = (float)sa[e];
            }
            __syncwarp();
            // lanes 0..5 own the columns of a, lanes 6..11 the columns of b
            const bool is_a = lane < 6 && lane < dims;
            const bool is_b = lane >= 6 && lane < 12 && lane - 6 < dims;
            const int col = lane < 6 ? lane : lane - 6;
            double det = 1.0;
            bool singular = false;
            for (int i = 0; i < dims; i++) {
                int k = i;
                for (int j = i + 1; j < dims; j++)
                    if (fabs(sa[j * 6 + i]) > fabs(sa[k * 6 + i])) k = j;
                if (fabs(sa[k * 6 + i]) < 2.220446049250313e-16 * 100) {
                    singular = true;
                    break;
                }
                __syncwarp();
                if (k != i) {
                    if ((is_a && col >= i) || is_b) {
                        double* m = is_a ? sa : sb;
                        const double tmp = m[i * 6 + col];
                        m[i * 6 + col] = m[k * 6 + col];
                        m[k * 6 + col] = tmp;
                    }
                    det = -det;
                    __syncwarp();
                }
                const double piv = sa[i * 6 + i];
                const double d = d_div(-1.0, piv);
                for (int r = i + 1; r < dims; r++) {
                    const double alpha = d_mul(sa[r * 6 + i], d);
                    if (is_a && col > i) sa[r * 6 + col] = d_add(sa[r * 6 + col], d_mul(alpha, sa[i * 6 + col]));
                    if (is_b) sb[r * 6 + col] = d_add(sb[r * 6 + col], d_mul(alpha, sb[i * 6 + col]));
                }
                det = d_mul(det, piv);
                __syncwarp();
            }
            if (singular) det = 0.0;
            if (det > 0) {  // inverse only written for det > 0 (aux_funs.cpp:104-111)
                if (is_b) {
                    for (int i = dims - 1; i >= 0; i--) {
                        double acc = sb[i * 6 + col];
                        for (int k = i + 1; k < dims; k++) acc = d_sub(acc, d_mul(sa[i * 6 + k], sb[k * 6 + col]));
                        sb[i * 6 + col] = d_div(acc, sa[i * 6 + i]);
                    }
                }
                __syncwarp();
                for (int e = lane; e < 36; e += 32) {
                    const int r = e / 6, c = e % 6;
                    if (r < dims && c <= r) s_cinv[(r * r + r) / 2 + c] = (float)sb[e];
                }
            }
            if (lane == 0) s_lu_state = det > 0 ? 0 : 2;
        }
        __syncthreads();
        VB_PHASE(0);
        if (s_lu_state == 2) {
            state = 2;
            break;
        }

        // E-step weights: hard truncation of the Mahalanobis distance (fit_robust_gaussian.cu:67-86)
        if constexpr (FAST6) {
            float mean[6], ci[21];
#pragma unroll
            for (int d = 0; d < 6; d++) mean[d] = s_mean[d];
#pragma unroll
            for (int k = 0; k < 21; k++) ci[k] = s_cinv[k];
            for (int li = t; li < n_local; li += kThreads) {
                if (S.gidx(li) >= N) continue;  // never read by the tree sums
                float sx[6], a[6];
#pragma unroll
                for (int d = 0; d < 6; d++) {
                    sx[d] = f_mul(S.local[(size_t)li * 6 + d], A.scale);
                    a[d] = f_sub(sx[d], mean[d]);
                }
                float z = 0.f;
#pragma unroll
                for (int d1 = 0; d1 < 6; d1++) {
                    float tmp = 0.f;
#pragma unroll
                    for (int d2 = 0; d2 < 6; d2++) {
                        const float c = (d1 >= d2) ? ci[(d1 * d1 + d1) / 2 + d2] : ci[(d2 * d2 + d2) / 2 + d1];
                        tmp = f_add(tmp, f_mul(c, a[d2]));
                    }
                    z = f_fma(tmp, a[d1], z);
                }
                z = __fsqrt_rn(z);
                const float w = z < A.trunc_sigma ? 1.f : 0.f;
                prod[li] = w;
#pragma unroll
                for (int d = 0; d < 6; d++) prod[(size_t)(1 + d) * n_local + li] = f_mul(w, sx[d]);
#pragma unroll
                for (int d1 = 0; d1 < 6; d1++)
#pragma unroll
                    for (int d2 = 0; d2 <= d1; d2++)
                        prod[(size_t)(7 + (d1 * d1 + d1) / 2 + d2) * n_local + li] = f_mul(f_mul(w, a[d1]), a[d2]);
            }
            __syncthreads();
            VB_PHASE(1);
            tree_level1(S, Q, X, cluster, [&](int q, int li) { return prod[(size_t)q * n_local + li]; });
        } else {
            for (int li = t; li < n_local; li += kThreads) {
                float wgt = 0.f;
                if (S.gidx(li) < N) {
                    float diff[kRobustMaxDims];
                    for (int d = 0; d < dims; d++) diff[d] = f_sub(f_mul(S.x(li, d), A.scale), s_mean[d]);
                    float z = 0.f;
                    for (int d1 = 0; d1 < dims; d1++) {
                        float tmp = 0.f;
                        for (int d2 = 0; d2 < dims; d2++) {
                            const float ci =
                                (d1 >= d2) ? s_cinv[(d1 * d1 + d1) / 2 + d2] : s_cinv[(d2 * d2 + d2) / 2 + d1];
                            tmp = f_add(tmp, f_mul(ci, diff[d2]));
                        }
                        z = f_fma(tmp, diff[d1], z);
                    }
                    z = __fsqrt_rn(z);
                    wgt = z < A.trunc_sigma ? 1.f : 0.f;
                    if (av)
                        for (int d = 0; d < dims; d++) av[(size_t)li * dims + d] = diff[d];
                }
                wv[li] = wgt;
            }
            __syncthreads();
            VB_PHASE(1);
            // weighted moments, reference tree order
            tree_level1(S, Q, X, cluster, [&](int q, int li) {
                const float w = wv[li];
                if (q == 0) return w;
                if (q <= dims) return f_mul(w, f_mul(S.x(li, q - 1), A.scale));
                const int d1 = s_d1[q - 1 - dims], d2 = s_d2[q - 1 - dims];
                const float a = av ? av[(size_t)li * dims + d1] : f_sub(f_mul(S.x(li, d1), A.scale), s_mean[d1]);
                const float b = av ? av[(size_t)li * dims + d2] : f_sub(f_mul(S.x(li, d2), A.scale), s_mean[d2]);
                return f_mul(f_mul(w, a), b);
            });
        }
        VB_PHASE(2);
        exchange_sync(X, cluster);
        VB_PHASE(3);
        tree_level2(S.NB, Q, X, sums);
        __syncthreads();
        VB_PHASE(4);
        {
            // host part of the reference iteration (fit_robust_gaussian.cu:209-246); the decision is evaluated by
            // every thread, the divisions of the M-step by one thread per moment
            const float prev_density = f_div(weight, (float)N);
            const float wsum = sums[0];
            weight = wsum;
            if (!isfinite(wsum)) {
                state = 2;
            } else if (fabsf(f_sub(f_div(wsum, (float)N), prev_density)) < A.epsilon) {
                state = 1;  // converged: keep the moments used in this E-step
            } else if (t < dims) {
                s_mean[t] = f_div(sums[1 + t], wsum);
            } else if (t < dims + cdims) {
                s_cov[t - dims] = f_div(sums[1 + t], wsum);
            }
        }
        __syncthreads();
        VB_PHASE(5);
        if (state != 0) break;
    }
#undef VB_PHASE
    if (rank == 0 && t == 0 && A.phase_cycles)
        for (int k = 0; k < 6; k++) A.phase_cycles[k] += ph[k];
    if (rank == 0 && t == 0) {
        out->reliable = (state != 2);
        out->used_iters = iter;
        out->density = f_div(weight, (float)N);
        for (int d = 0; d < dims; d++) out->mean[d] = s_mean[d];
        for (int k = 0; k < cdims; k++) out->covar[k] = s_cov[k];
    }
}

// shared-memory plan for a pool of at most n elements: weights + (if it fits) the slice of every CTA
// fast_q > 0: also try to place the [fast_q][n_local] product rows of the FAST6 kernels (dims == 6 only)
void smem_plan(int n, int dims, int& slice_in_smem, size_t& bytes, int* centred_in_smem, int fast_q, bool& fast6) {
    const int NB = (n + 511) / 512;
    const size_t nlb = (size_t)(NB + kCluster - 1) / kCluster;
    const size_t wb = nlb * 512 * sizeof(float);
    const size_t pb = nlb * 512 * dims * sizeof(float);
    slice_in_smem = (wb + pb <= kSmemBudget);
    bytes = wb + (slice_in_smem ? pb : 0);
    fast6 = dims == 6 && fast_q > 0 && slice_in_smem && bytes + wb * fast_q <= kSmemBudget;
    if (fast6) {
        bytes += wb * fast_q;
        if (centred_in_smem) *centred_in_smem = 0;
        return;
    }
    if (centred_in_smem) {
        *centred_in_smem = (bytes + pb <= kSmemBudget);
        if (*centred_in_smem) bytes += pb;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int PoseMode::init() {
    if (stream) return 0;
    VB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    VB_CUDA(cudaMalloc((void**)&d_result, kMaxFrames * sizeof(MeanshiftResult)));
    VB_CUDA(cudaMallocHost((void**)&h_result, kMaxFrames * sizeof(MeanshiftResult)));
    VB_CUDA(cudaMalloc((void**)&d_partials, (size_t)64 * kMaxTreeBlocks * sizeof(float)));
    VB_CUDA(cudaMalloc((void**)&d_rg_sums, 256 + kCluster * sizeof(int)));
    VB_CUDA(cudaMallocHost((void**)&h_rg_sums, 256));
    if (getenv("VB_POSE_MODE_PHASES")) {
        VB_CUDA(cudaMalloc((void**)&d_phase_cycles, 24 * sizeof(long long)));
        VB_CUDA(cudaMemset(d_phase_cycles, 0, 24 * sizeof(long long)));
    }
    VB_CUDA(cudaFuncSetAttribute(k_meanshift<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    VB_CUDA(cudaFuncSetAttribute(k_meanshift<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    VB_CUDA(cudaFuncSetAttribute(k_robust_fit<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    VB_CUDA(cudaFuncSetAttribute(k_robust_fit<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    return 0;
}

static int run_meanshift(PoseMode& M, MeanshiftArgs& A, int n_plan, float* h_io_mean, float* h_o_confidence,
                         int* used_iters) {
    if (n_plan > 512 * 512) return (int)cudaErrorInvalidValue;
    size_t smem_bytes;
    bool fast6;
    smem_plan(n_plan, A.dims, A.slice_in_smem, smem_bytes, nullptr, A.trial_only ? 1 : A.dims + 1, fast6);
    A.src.cta_counts = (int*)((char*)M.d_rg_sums + 256);
    A.phase_cycles = M.d_phase_cycles;
    if (fast6)
        k_meanshift<true><<<kCluster, kThreads, smem_bytes, M.stream>>>(A, M.d_partials, M.d_result);
    else
        k_meanshift<false><<<kCluster, kThreads, smem_bytes, M.stream>>>(A, M.d_partials, M.d_result);
    VB_RETURN_IF_CUDA_ERROR();
    VB_CUDA(cudaMemcpyAsync(M.h_result, M.d_result, sizeof(MeanshiftResult), cudaMemcpyDeviceToHost, M.stream));
    VB_CUDA(cudaStreamSynchronize(M.stream));
    if (M.prof) {
        if (A.trial_only)
            M.prof->meanshift_trials++;
        else
            M.prof->meanshift_runs++, M.prof->meanshift_iters += M.h_result->used_iters;
    }
    if (!A.trial_only && M.h_result->used_iters > 0) {
        if (h_o_confidence) *h_o_confidence = M.h_result->confidence;
        if (used_iters) *used_iters = M.h_result->used_iters;
        for (int d = 0; d < A.dims; d++) h_io_mean[d] = M.h_result->mean[d];
    }
    return 0;
}

int PoseMode::meanshift(const float* d_space, const float* h_space_for_init, const int* d_n, int n_host, int dims,
                        float kernel_var, float* h_io_mean, float* h_o_confidence, int* used_iters,
                        bool use_external_init_mean, float epsilon, int max_iters, int max_init_trials,
                        float good_init_confidence, int n_capacity) {
    (void)h_space_for_init;
    if (int e = init()) return e;
    if (dims > kMeanshiftMaxDims) return (int)cudaErrorInvalidValue;
    MeanshiftArgs A;
    memset(&A, 0, sizeof(A));
    for (int d = 0; d < dims; d++) A.io_mean[d] = h_io_mean[d];
    A.dims = dims, A.kernel_var = kernel_var, A.epsilon = epsilon, A.max_iters = max_iters;
    A.center_idx = -1, A.trial_only = 0;
    A.src.space = d_space, A.src.d_n = d_n, A.src.n_host = n_host;

    int N = n_host;
    if (!use_external_init_mean) {
        // start from the best of <= max_init_trials samples drawn with the host libc rand() (meanshift.cu:73-97;
        // the unseeded process-wide stream is part of the reference behaviour)
        if (d_n) {
            VB_CUDA(cudaMemcpyAsync(&h_result->n, d_n, sizeof(int), cudaMemcpyDeviceToHost, stream));
            VB_CUDA(cudaStreamSynchronize(stream));
            N = h_result->n;
        }
        if (N <= 0) return (int)cudaErrorInvalidValue;
        A.src.d_n = nullptr, A.src.n_host = N;
        d_n = nullptr;
        if (!rnd) return (int)cudaErrorInvalidValue;
        if (max_init_trials > 0 && max_init_trials <= kMaxTrialBatch && rnd->snapshot()) {
            // all trials + the selection loop + the iteration in one launch; the libc stream is rewound to the
            // number of draws the reference's early-exit loop would have consumed (libc_rand.h)
            A.n_trials = max_init_trials;
            A.good_init_confidence = good_init_confidence;
            for (int trial = 0; trial < max_init_trials; trial++) A.trial_idx[trial] = rnd->next() % N;
            if (used_iters) *used_iters = 0;
            if (int e = run_meanshift(*this, A, N, h_io_mean, h_o_confidence, used_iters)) return e;
            const int used = h_result->trials_used;
            if (prof) prof->meanshift_trials += used;
            if (used < max_init_trials) {
                rnd->rewind();
                for (int trial = 0; trial < used; trial++) (void)rnd->next();
            }
            return 0;
        }
        float best_conf = 0;
        int best_idx = -1;
        for (int trial = 0; trial < max_init_trials; trial++) {
            const int idx_rand = (rnd->next() % N);
            MeanshiftArgs T = A;
            T.center_idx = idx_rand, T.trial_only = 1;
            if (int e = run_meanshift(*this, T, N, nullptr, nullptr, nullptr)) return e;
            if (h_result->weight_sum > best_conf) {
                best_conf = h_result->weight_sum;
                best_idx = idx_rand;
            }
            if (best_conf > good_init_confidence * N) break;
        }
        A.center_idx = best_idx < 0 ? 0 : best_idx;
    }
    if (used_iters) *used_iters = 0;
    return run_meanshift(*this, A, d_n ? n_capacity : N, h_io_mean, h_o_confidence, used_iters);
}

int PoseMode::meanshift_from_hypotheses(const float* d_rvecs, const float* d_tvecs, int n_poses, float rvec_scale,
                                        float* d_pool, int* d_used, int dims, float kernel_var, float* h_io_mean,
                                        float* h_o_confidence, int* used_iters, float epsilon, int max_iters,
                                        const int* d_aux_count) {
    if (int e = init()) return e;
    MeanshiftArgs A;
    memset(&A, 0, sizeof(A));
    for (int d = 0; d < dims; d++) A.io_mean[d] = h_io_mean[d];
    A.dims = dims, A.kernel_var = kernel_var, A.epsilon = epsilon, A.max_iters = max_iters;
    A.center_idx = -1, A.trial_only = 0;
    A.src.rvecs = d_rvecs, A.src.tvecs = d_tvecs, A.src.n_poses = n_poses, A.src.rvec_scale = rvec_scale;
    A.src.pool_out = d_pool, A.src.used_out = d_used;
    A.src.aux_count = d_aux_count;
    if (used_iters) *used_iters = 0;
    return run_meanshift(*this, A, n_poses, h_io_mean, h_o_confidence, used_iters);
}

int PoseMode::enqueue_from_hypotheses(int slot, const float* d_rvecs, const float* d_tvecs, int n_poses,
                                      float rvec_scale, float* d_pool, int* d_used, int dims, float kernel_var,
                                      const float* h_init_mean, float epsilon, int max_iters,
                                      const int* d_aux_count, const PoseTail& tail) {
    if (int e = init()) return e;
    if (slot < 0 || slot >= kMaxFrames || n_poses > 512 * 512) return (int)cudaErrorInvalidValue;
    MeanshiftArgs A;
    memset(&A, 0, sizeof(A));
    for (int d = 0; d < dims; d++) A.io_mean[d] = h_init_mean[d];
    A.dims = dims, A.kernel_var = kernel_var, A.epsilon = epsilon, A.max_iters = max_iters;
    A.center_idx = -1, A.trial_only = 0;
    A.src.rvecs = d_rvecs, A.src.tvecs = d_tvecs, A.src.n_poses = n_poses, A.src.rvec_scale = rvec_scale;
    A.src.pool_out = d_pool, A.src.used_out = d_used;
    A.src.aux_count = d_aux_count;
    A.src.cta_counts = (int*)((char*)d_rg_sums + 256);
    A.tail = tail;
    A.phase_cycles = d_phase_cycles;
    size_t smem_bytes;
    bool fast6;
    smem_plan(n_poses, dims, A.slice_in_smem, smem_bytes, nullptr, dims + 1, fast6);
    if (fast6)
        k_meanshift<true><<<kCluster, kThreads, smem_bytes, stream>>>(A, d_partials, d_result + slot);
    else
        k_meanshift<false><<<kCluster, kThreads, smem_bytes, stream>>>(A, d_partials, d_result + slot);
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

int PoseMode::fetch_results(int n_slots) {
    if (n_slots <= 0) return 0;
    VB_CUDA(cudaMemcpyAsync(h_result, d_result, (size_t)n_slots * sizeof(MeanshiftResult), cudaMemcpyDeviceToHost, stream));
    VB_CUDA(cudaStreamSynchronize(stream));
    if (prof)
        for (int i = 0; i < n_slots; i++) prof->meanshift_runs++, prof->meanshift_iters += h_result[i].used_iters;
    return 0;
}

int PoseMode::fit_robust_gaussian(const float* d_space, int N, int dims, float scale, float* h_io_mean,
                                  float* h_io_covar, float trunc_sigma, float covar_reg_lambda, float* h_o_density,
                                  int* used_iters, float epsilon, int max_iters) {
    if (int e = init()) return e;
    if (dims > kRobustMaxDims) throw;  // reference: fit_robust_gaussian.cu:107-108
    if (N > 512 * 512 || N <= 0) return (int)cudaErrorInvalidValue;
    RobustArgs A;
    memset(&A, 0, sizeof(A));
    for (int d = 0; d < dims; d++) A.mean[d] = h_io_mean[d];
    for (int d1 = 0; d1 < dims; d1++)
        for (int d2 = 0; d2 <= d1; d2++) A.covar[(d1 * d1 + d1) / 2 + d2] = h_io_covar[d1 * dims + d2];
    A.trunc_sigma = trunc_sigma, A.scale = scale, A.covar_reg_lambda = covar_reg_lambda, A.epsilon = epsilon;
    A.N = N, A.dims = dims, A.max_iters = max_iters, A.space = d_space;
    size_t smem_bytes;
    bool fast6;
    smem_plan(N, dims, A.slice_in_smem, smem_bytes, &A.centred_in_smem, 1 + dims + (dims * dims + dims) / 2, fast6);
    A.phase_cycles = d_phase_cycles;
    if (used_iters) *used_iters = 0;

    RobustResult* d_res = (RobustResult*)d_rg_sums;
    RobustResult* h_res = (RobustResult*)h_rg_sums;
    if (fast6)
        k_robust_fit<true><<<kCluster, kThreads, smem_bytes, stream>>>(A, d_partials, d_res);
    else
        k_robust_fit<false><<<kCluster, kThreads, smem_bytes, stream>>>(A, d_partials, d_res);
    VB_RETURN_IF_CUDA_ERROR();
    VB_CUDA(cudaMemcpyAsync(h_res, d_res, sizeof(RobustResult), cudaMemcpyDeviceToHost, stream));
    VB_CUDA(cudaStreamSynchronize(stream));
    if (prof) prof->robust_runs++, prof->robust_iters += h_res->used_iters;

    if (h_res->reliable) {
        if (h_o_density) *h_o_density = h_res->density;
        if (used_iters) *used_iters = h_res->used_iters;
        for (int d1 = 0; d1 < dims; d1++)
            for (int d2 = 0; d2 <= d1; d2++) {
                h_io_covar[d1 * dims + d2] = h_res->covar[(d1 * d1 + d1) / 2 + d2];
                h_io_covar[d2 * dims + d1] = h_io_covar[d1 * dims + d2];
            }
        for (int d = 0; d < dims; d++) h_io_mean[d] = h_res->mean[d];
    }
    return h_res->reliable ? 0 : 1;  // cudaSuccess / !cudaSuccess (fit_robust_gaussian.cu:281-284)
}

}  // namespace vb
