// Mean-shift and robust Gaussian fit kernels for sm_100a.  See pose_mode.cuh.
//
// Both run as ONE launch of ONE thread-block cluster (8 CTAs x 512 threads, distributed over 8 SMs): the pose
// pool (<= 8192 x 6 floats) is split by 512-element tree blocks over the CTAs and staged in their shared memory;
// every iteration of the fixed-point loop runs on the device — weights, the reference-ordered tree sums
// (tree_sum.cuh; level-1 partials are exchanged through L2 with one hardware cluster barrier per iteration,
// double buffered), the host-side arithmetic of the reference's loop (mean update / displacement test; 6x6 FP64
// LU, inverse and shrinkage for the robust fit — evaluated redundantly and identically by every CTA) and the
// convergence test.  One launch replaces the reference's up to 100 x (kernel + 2..3 multi-pass reductions +
// 2..4 blocking copies) per call.
#include "pose_mode.cuh"
#include "residual_model.cuh"
#include "tree_sum.cuh"
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
    int* cta_counts;  // [kCluster] scratch
};

struct MeanshiftArgs {
    float io_mean[kMeanshiftMaxDims];  // caller's mean: first displacement is measured against it (Q13)
    int center_idx;                    // >= 0: start from space[center_idx]; < 0: start from io_mean
    int trial_only;                    // 1: only the kernel-weight sum around the start point
    int dims;
    float kernel_var, epsilon;
    int max_iters;
    int slice_in_smem;
    PoolSource src;
};

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

// level 1: one warp per (local tree block, quantity); level 2: one warp per quantity (every CTA, redundantly)
template <class Val>
__device__ __forceinline__ void tree_level1(const Slice& S, int Q, float* partials, Val val) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int task = warp; task < S.nlb * Q; task += kWarps) {
        const int q = task / S.nlb, lb = task % S.nlb;
        const int b = lb * kCluster + S.rank;
        if (b >= S.NB) continue;
        const int count = min(512, S.N - b * 512);
        float v = tree_sum_512([&](int i) { return val(q, lb * 512 + i); }, count, lane);
        if (lane == 0) {
            if (S.N == 1) v = val(q, 0);  // the reference launches no reduction at all for a single element
            partials[q * kMaxTreeBlocks + b] = v;
        }
    }
}
__device__ __forceinline__ void tree_level2(int NB, int Q, const float* partials, float* sums) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const volatile float* pv = partials;
    for (int q = warp; q < Q; q += kWarps) {
        const volatile float* p = pv + q * kMaxTreeBlocks;
        const float v = (NB == 1) ? p[0] : tree_sum_512([&](int i) { return (float)p[i]; }, NB, lane);
        if (lane == 0) sums[q] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// mean-shift (reference: meanshift.cu:12-31 weights, :99-134 host loop)
// ------------------------------------------------------------------------------------------------
__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
    k_meanshift(const MeanshiftArgs A, float* partials_g, MeanshiftResult* out) {
    extern __shared__ float smem[];
    __shared__ float c_mean[kMeanshiftMaxDims];
    __shared__ float io_mean[kMeanshiftMaxDims];
    __shared__ float sums[kMeanshiftMaxDims + 1];
    __shared__ int done;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int dims = A.dims;
    const int Q = A.trial_only ? 1 : dims + 1;

    int N;
    const float* space_g = A.src.space;
    if (A.src.rvecs) {
        N = build_pool(A.src, cluster, rank);
        space_g = A.src.pool_out;
    } else {
        N = A.src.d_n ? *A.src.d_n : A.src.n_host;
    }
    if (N <= 0) {
        if (rank == 0 && threadIdx.x == 0) out->used_iters = 0, out->n = N, out->weight_sum = 0.f, out->confidence = 0.f;
        return;
    }
    Slice S;
    S.N = N, S.NB = (N + 511) / 512, S.rank = rank, S.dims = dims, S.global = space_g;
    S.nlb = (S.NB - rank + kCluster - 1) / kCluster;
    if (S.nlb < 0) S.nlb = 0;
    float* wv = smem;  // weights of the local slice
    stage_slice(S, smem + (size_t)S.nlb * 512, A.slice_in_smem != 0);
    if (threadIdx.x < dims) {
        io_mean[threadIdx.x] = A.io_mean[threadIdx.x];
        c_mean[threadIdx.x] = (A.center_idx >= 0)
                                  ? ((const volatile float*)space_g)[(size_t)A.center_idx * dims + threadIdx.x]
                                  : A.io_mean[threadIdx.x];
    }
    if (threadIdx.x == 0) done = 0;
    __syncthreads();

    const float two_var = f_add(A.kernel_var, A.kernel_var);  // 2*kernel_var
    const int n_iters = A.trial_only ? 1 : A.max_iters;
    int used_iters = 0;
    float confidence = 0.f, wsum_last = 0.f;

    for (int iter = 0; iter < n_iters; iter++) {
        float* partials = partials_g + (size_t)(iter & 1) * 32 * kMaxTreeBlocks;
        // weights w_i = exp(-|x_i - mu|^2 / (2 var)) of the local slice
        for (int li = threadIdx.x; li < S.nlb * 512; li += kThreads) {
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
        tree_level1(S, Q, partials, [&](int q, int li) { return q == 0 ? wv[li] : f_mul(wv[li], S.x(li, q - 1)); });
        __threadfence();
        cluster.sync();
        tree_level2(S.NB, Q, partials, sums);
        __syncthreads();
        if (threadIdx.x == 0) {
            const float wsum = sums[0];
            wsum_last = wsum;
            if (!A.trial_only) {
                // host part of the reference iteration (meanshift.cu:112-133)
                float mean_new[kMeanshiftMaxDims];
                for (int d = 0; d < dims; d++) mean_new[d] = f_div(sums[d + 1], wsum);
                confidence = f_div(wsum, (float)N);
                used_iters = iter + 1;
                float disp = 0.f;
                for (int d = 0; d < dims; d++) {
                    const float df = f_sub(io_mean[d], mean_new[d]);
                    disp = f_add(disp, f_mul(df, df));
                }
                disp = __fsqrt_rn(disp);
                for (int d = 0; d < dims; d++) io_mean[d] = mean_new[d], c_mean[d] = mean_new[d];
                if (disp < A.epsilon) done = 1;
            }
        }
        __syncthreads();
        if (done) break;
    }
    if (rank == 0 && threadIdx.x == 0) {
        for (int d = 0; d < dims; d++) out->mean[d] = io_mean[d];
        out->confidence = confidence;
        out->weight_sum = wsum_last;
        out->used_iters = used_iters;
        out->n = N;
    }
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
    int slice_in_smem;
    const float* space;
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

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
    k_robust_fit(const RobustArgs A, float* partials_g, RobustResult* out) {
    extern __shared__ float smem[];
    __shared__ float s_mean[kRobustMaxDims];
    __shared__ float s_cov[21], s_cinv[21];
    __shared__ float sums[28];
    __shared__ int s_state;  // 0 continue, 1 converged/reliable, 2 unreliable
    __shared__ float s_weight;
    // 6x6 FP64 scratch of the covariance step: LU with partial pivoting run by the 36 threads (r,c) of the
    // augmented matrix [a | b] in shared memory — the same operations on the same elements as the sequential host
    // code it mirrors (aux shim / cv::Matx66d), only issued in parallel
    __shared__ double sa[36], sb[36], s_inv[36];
    __shared__ double s_det;
    __shared__ int s_piv;
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int N = A.N, dims = A.dims;
    const int cdims = (dims * dims + dims) / 2;
    const int Q = 1 + dims + cdims;

    Slice S;
    S.N = N, S.NB = (N + 511) / 512, S.rank = rank, S.dims = dims, S.global = A.space;
    S.nlb = (S.NB - rank + kCluster - 1) / kCluster;
    if (S.nlb < 0) S.nlb = 0;
    float* wv = smem;
    stage_slice(S, smem + (size_t)S.nlb * 512, A.slice_in_smem != 0);
    const int t = threadIdx.x;
    if (t == 0) {
        for (int d = 0; d < dims; d++) s_mean[d] = A.mean[d];
        for (int k = 0; k < cdims; k++) s_cov[k] = A.covar[k];
        s_weight = 0.f;
        s_state = 0;
    }
    const int mr = t / 6, mc = t % 6;
    const bool in_mat = t < 36 && mr < dims && mc < dims;
    if (t < 36) s_inv[t] = 0.0;
    __syncthreads();

    int iter = 0;
    for (iter = 0; iter < A.max_iters; iter++) {
        float* partials = partials_g + (size_t)(iter & 1) * 32 * kMaxTreeBlocks;
        if (t == 0) {
            // half -> full (double); shrink towards tr/n * I from the 2nd iteration on
            for (int d1 = 0; d1 < dims; d1++)
                for (int d2 = 0; d2 <= d1; d2++) {
                    sa[d1 * 6 + d2] = (double)s_cov[(d1 * d1 + d1) / 2 + d2];
                    if (d1 != d2) sa[d2 * 6 + d1] = sa[d1 * 6 + d2];
                }
            if (iter > 0 && A.covar_reg_lambda > 0) {
                const double lambda = (double)A.covar_reg_lambda;
                double tr = 0;
                for (int i = 0; i < dims; i++) tr = d_add(tr, sa[i * 6 + i]);
                const double m = d_div(tr, (double)dims);
                const double lm = d_mul(lambda, m), oml = d_sub(1.0, lambda);
                for (int i = 0; i < dims; i++)
                    for (int j = 0; j < dims; j++)
                        sa[i * 6 + j] = d_add(d_mul(lm, i == j ? 1.0 : 0.0), d_mul(oml, sa[i * 6 + j]));
            }
            // the regularised covariance is the one reported on convergence (fit_robust_gaussian.cu:203)
            for (int d1 = 0; d1 < dims; d1++)
                for (int d2 = 0; d2 <= d1; d2++) s_cov[(d1 * d1 + d1) / 2 + d2] = (float)sa[d1 * 6 + d2];
            s_det = 1.0;
        }
        if (in_mat) sb[t] = (mr == mc) ? 1.0 : 0.0;
        __syncthreads();
        bool singular = false;
        for (int i = 0; i < dims; i++) {
            if (t == 0) {
                int k = i;
                for (int j = i + 1; j < dims; j++)
                    if (fabs(sa[j * 6 + i]) > fabs(sa[k * 6 + i])) k = j;
                s_piv = (fabs(sa[k * 6 + i]) < 2.220446049250313e-16 * 100) ? -1 : k;
            }
            __syncthreads();
            const int k = s_piv;
            if (k < 0) {
                singular = true;
                break;
            }
            if (k != i) {
                if (in_mat && mr == i) {
                    if (mc >= i) {
                        const double tmp = sa[i * 6 + mc];
                        sa[i * 6 + mc] = sa[k * 6 + mc];
                        sa[k * 6 + mc] = tmp;
                    }
                    const double tmp = sb[i * 6 + mc];
                    sb[i * 6 + mc] = sb[k * 6 + mc];
                    sb[k * 6 + mc] = tmp;
                }
                if (t == 0) s_det = -s_det;
                __syncthreads();
            }
            double na = 0, nb = 0;
            const bool upd = in_mat && mr > i;
            if (upd) {
                const double d = d_div(-1.0, sa[i * 6 + i]);
                const double alpha = d_mul(sa[mr * 6 + i], d);
                na = (mc > i) ? d_add(sa[mr * 6 + mc], d_mul(alpha, sa[i * 6 + mc])) : sa[mr * 6 + mc];
                nb = d_add(sb[mr * 6 + mc], d_mul(alpha, sb[i * 6 + mc]));
            }
            __syncthreads();
            if (upd) sa[mr * 6 + mc] = na, sb[mr * 6 + mc] = nb;
            if (t == 0) s_det = d_mul(s_det, sa[i * 6 + i]);
            __syncthreads();
        }
        const double det = singular ? 0.0 : s_det;
        if (det > 0) {  // inverse only written for det > 0 (aux_funs.cpp:104-111)
            for (int i = dims - 1; i >= 0; i--) {
                if (t < dims) {
                    double acc = sb[i * 6 + t];
                    for (int k = i + 1; k < dims; k++) acc = d_sub(acc, d_mul(sa[i * 6 + k], sb[k * 6 + t]));
                    sb[i * 6 + t] = d_div(acc, sa[i * 6 + i]);
                }
                __syncthreads();
            }
            if (in_mat) s_inv[t] = sb[t];
        }
        __syncthreads();
        if (t == 0) {
            if (det <= 0) {
                s_state = 2;
            } else {
                for (int d1 = 0; d1 < dims; d1++)
                    for (int d2 = 0; d2 <= d1; d2++) s_cinv[(d1 * d1 + d1) / 2 + d2] = (float)s_inv[d1 * 6 + d2];
            }
        }
        __syncthreads();
        if (s_state == 2) break;

        // E-step weights: hard truncation of the Mahalanobis distance (fit_robust_gaussian.cu:67-86)
        for (int li = t; li < S.nlb * 512; li += kThreads) {
            float wgt = 0.f;
            if (S.gidx(li) < N) {
                float diff[kRobustMaxDims];
                for (int d = 0; d < dims; d++) diff[d] = f_sub(f_mul(S.x(li, d), A.scale), s_mean[d]);
                float z = 0.f;
                for (int d1 = 0; d1 < dims; d1++) {
                    float tmp = 0.f;
                    for (int d2 = 0; d2 < dims; d2++) {
                        const float ci = (d1 >= d2) ? s_cinv[(d1 * d1 + d1) / 2 + d2] : s_cinv[(d2 * d2 + d2) / 2 + d1];
                        tmp = f_add(tmp, f_mul(ci, diff[d2]));
                    }
                    z = f_fma(tmp, diff[d1], z);
                }
                z = __fsqrt_rn(z);
                wgt = z < A.trunc_sigma ? 1.f : 0.f;
            }
            wv[li] = wgt;
        }
        __syncthreads();
        // weighted moments, reference tree order
        tree_level1(S, Q, partials, [&](int q, int li) {
            const float w = wv[li];
            if (q == 0) return w;
            if (q <= dims) return f_mul(w, f_mul(S.x(li, q - 1), A.scale));
            int k = q - 1 - dims, d1 = 0;
            while ((d1 + 1) * (d1 + 2) / 2 <= k) d1++;
            const int d2 = k - (d1 * d1 + d1) / 2;
            const float a = f_sub(f_mul(S.x(li, d1), A.scale), s_mean[d1]);
            const float b = f_sub(f_mul(S.x(li, d2), A.scale), s_mean[d2]);
            return f_mul(f_mul(w, a), b);
        });
        __threadfence();
        cluster.sync();
        tree_level2(S.NB, Q, partials, sums);
        __syncthreads();
        if (t == 0) {
            // host part of the reference iteration (fit_robust_gaussian.cu:209-246)
            const float prev_density = f_div(s_weight, (float)N);
            const float wsum = sums[0];
            s_weight = wsum;
            if (!isfinite(wsum)) {
                s_state = 2;
            } else if (fabsf(f_sub(f_div(wsum, (float)N), prev_density)) < A.epsilon) {
                s_state = 1;  // converged: keep the moments used in this E-step
            } else {
                for (int d = 0; d < dims; d++) s_mean[d] = f_div(sums[1 + d], wsum);
                for (int k = 0; k < cdims; k++) s_cov[k] = f_div(sums[1 + dims + k], wsum);
            }
        }
        __syncthreads();
        if (s_state != 0) break;
    }
    if (rank == 0 && t == 0) {
        out->reliable = (s_state != 2);
        out->used_iters = iter;
        out->density = f_div(s_weight, (float)N);
        for (int d = 0; d < dims; d++) out->mean[d] = s_mean[d];
        for (int k = 0; k < cdims; k++) out->covar[k] = s_cov[k];
    }
}

// shared-memory plan for a pool of at most n elements: weights + (if it fits) the slice of every CTA
void smem_plan(int n, int dims, int& slice_in_smem, size_t& bytes) {
    const int NB = (n + 511) / 512;
    const size_t nlb = (size_t)(NB + kCluster - 1) / kCluster;
    const size_t wb = nlb * 512 * sizeof(float);
    const size_t pb = nlb * 512 * dims * sizeof(float);
    slice_in_smem = (wb + pb <= kSmemBudget);
    bytes = wb + (slice_in_smem ? pb : 0);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int PoseMode::init() {
    if (stream) return 0;
    VB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    VB_CUDA(cudaMalloc((void**)&d_result, sizeof(MeanshiftResult)));
    VB_CUDA(cudaMallocHost((void**)&h_result, sizeof(MeanshiftResult)));
    VB_CUDA(cudaMalloc((void**)&d_partials, (size_t)64 * kMaxTreeBlocks * sizeof(float)));
    VB_CUDA(cudaMalloc((void**)&d_rg_sums, 256 + kCluster * sizeof(int)));
    VB_CUDA(cudaMallocHost((void**)&h_rg_sums, 256));
    VB_CUDA(cudaFuncSetAttribute(k_meanshift, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    VB_CUDA(cudaFuncSetAttribute(k_robust_fit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
    return 0;
}

static int run_meanshift(PoseMode& M, MeanshiftArgs& A, int n_plan, float* h_io_mean, float* h_o_confidence,
                         int* used_iters) {
    if (n_plan > 512 * 512) return (int)cudaErrorInvalidValue;
    size_t smem_bytes;
    smem_plan(n_plan, A.dims, A.slice_in_smem, smem_bytes);
    A.src.cta_counts = (int*)((char*)M.d_rg_sums + 256);
    k_meanshift<<<kCluster, kThreads, smem_bytes, M.stream>>>(A, M.d_partials, M.d_result);
    VB_RETURN_IF_CUDA_ERROR();
    VB_CUDA(cudaMemcpyAsync(M.h_result, M.d_result, sizeof(MeanshiftResult), cudaMemcpyDeviceToHost, M.stream));
    VB_CUDA(cudaStreamSynchronize(M.stream));
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
        float best_conf = 0;
        int best_idx = -1;
        for (int trial = 0; trial < max_init_trials; trial++) {
            const int idx_rand = (rand() % N);
            MeanshiftArgs T = A;
            T.center_idx = idx_rand, T.trial_only = 1;
            T.src.d_n = nullptr, T.src.n_host = N;
            if (int e = run_meanshift(*this, T, N, nullptr, nullptr, nullptr)) return e;
            if (h_result->weight_sum > best_conf) {
                best_conf = h_result->weight_sum;
                best_idx = idx_rand;
            }
            if (best_conf > good_init_confidence * N) break;
        }
        A.center_idx = best_idx < 0 ? 0 : best_idx;
        A.src.d_n = nullptr, A.src.n_host = N;
        d_n = nullptr;
    }
    if (used_iters) *used_iters = 0;
    return run_meanshift(*this, A, d_n ? n_capacity : N, h_io_mean, h_o_confidence, used_iters);
}

int PoseMode::meanshift_from_hypotheses(const float* d_rvecs, const float* d_tvecs, int n_poses, float rvec_scale,
                                        float* d_pool, int* d_used, int dims, float kernel_var, float* h_io_mean,
                                        float* h_o_confidence, int* used_iters, float epsilon, int max_iters) {
    if (int e = init()) return e;
    MeanshiftArgs A;
    memset(&A, 0, sizeof(A));
    for (int d = 0; d < dims; d++) A.io_mean[d] = h_io_mean[d];
    A.dims = dims, A.kernel_var = kernel_var, A.epsilon = epsilon, A.max_iters = max_iters;
    A.center_idx = -1, A.trial_only = 0;
    A.src.rvecs = d_rvecs, A.src.tvecs = d_tvecs, A.src.n_poses = n_poses, A.src.rvec_scale = rvec_scale;
    A.src.pool_out = d_pool, A.src.used_out = d_used;
    if (used_iters) *used_iters = 0;
    return run_meanshift(*this, A, n_poses, h_io_mean, h_o_confidence, used_iters);
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
    smem_plan(N, dims, A.slice_in_smem, smem_bytes);
    if (used_iters) *used_iters = 0;

    RobustResult* d_res = (RobustResult*)d_rg_sums;
    RobustResult* h_res = (RobustResult*)h_rg_sums;
    k_robust_fit<<<kCluster, kThreads, smem_bytes, stream>>>(A, d_partials, d_res);
    VB_RETURN_IF_CUDA_ERROR();
    VB_CUDA(cudaMemcpyAsync(h_res, d_res, sizeof(RobustResult), cudaMemcpyDeviceToHost, stream));
    VB_CUDA(cudaStreamSynchronize(stream));

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

PoseMode& global_pose_mode() {
    static PoseMode inst;
    return inst;
}

}  // namespace vb
