// Mean-shift and robust Gaussian fit kernels for sm_100a.  See pose_mode.cuh.
#include "pose_mode.cuh"
#include "residual_model.cuh"
#include "small_linalg.h"
#include "tree_sum.cuh"
#include <cmath>
#include <cstdlib>

namespace vb {

namespace {

constexpr int kMsThreads = 512;           // 16 warps: one 512-element tree block per warp per round
constexpr int kMsWarps = kMsThreads / 32;
constexpr int kMaxTreeBlocks = 512;       // pool size limit 512*512 (two tree levels)

struct MeanshiftArgs {
    float io_mean[kMeanshiftMaxDims];  // caller's mean: first displacement is measured against it (Q13)
    int center_idx;                    // >= 0: start from space[center_idx]; < 0: start from io_mean
    int trial_only;                    // 1: only the kernel-weight sum around the start point
    int n_host, dims;
    float kernel_var, epsilon;
    int max_iters;
};

// weight of sample i around c_mean (reference: meanshift.cu:20-25)
__device__ __forceinline__ float ms_weight(const float* __restrict__ space, int i, int dims, const float* c_mean,
                                           float two_var) {
    float l2 = 0.f;
    for (int d = 0; d < dims; d++) {
        const float diff = f_sub(space[(size_t)i * dims + d], c_mean[d]);
        l2 = f_fma(diff, diff, l2);
    }
    return expf(f_div(-l2, two_var));
}

__global__ void __launch_bounds__(kMsThreads)
    k_meanshift(const float* __restrict__ space, const int* d_n, const MeanshiftArgs A, float* partials,
                MeanshiftResult* out) {
    __shared__ float c_mean[kMeanshiftMaxDims];
    __shared__ float io_mean[kMeanshiftMaxDims];
    __shared__ float sums[kMeanshiftMaxDims + 1];
    __shared__ int done;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = d_n ? *d_n : A.n_host;
    const int dims = A.dims;
    const int Q = dims + 1;
    if (threadIdx.x < dims) {
        io_mean[threadIdx.x] = A.io_mean[threadIdx.x];
        c_mean[threadIdx.x] =
            (A.center_idx >= 0 && N > 0) ? space[(size_t)A.center_idx * dims + threadIdx.x] : A.io_mean[threadIdx.x];
    }
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    if (N <= 0) {
        if (threadIdx.x == 0) out->used_iters = 0, out->n = N, out->weight_sum = 0.f, out->confidence = 0.f;
        return;
    }
    const int NB = (N + 511) / 512;
    const float two_var = f_add(A.kernel_var, A.kernel_var);  // 2*kernel_var
    const int n_iters = A.trial_only ? 1 : A.max_iters;
    int used = 0;
    float confidence = 0.f, wsum_last = 0.f;

    for (int iter = 0; iter < n_iters; iter++) {
        // level 1: one warp per 512-block, all Q quantities
        for (int b = warp; b < NB; b += kMsWarps) {
            const int base = b * 512;
            const int count = min(512, N - base);
            float wgt[16];
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int t = lane + 32 * j;
                wgt[j] = (t < count) ? ms_weight(space, base + t, dims, c_mean, two_var) : 0.f;
            }
            for (int q = 0; q < (A.trial_only ? 1 : Q); q++) {
                float a[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int t = lane + 32 * j;
                    float v = 0.f;
                    if (t < count) {
                        v = (q == 0) ? wgt[j] : f_mul(wgt[j], space[(size_t)(base + t) * dims + q - 1]);
                        if (t + 256 < count) {
                            const float v2 = (q == 0) ? wgt[j + 8]
                                                      : f_mul(wgt[j + 8], space[(size_t)(base + t + 256) * dims + q - 1]);
                            v = f_add(v, v2);
                        }
                    }
                    a[j] = v;
                }
#pragma unroll
                for (int j = 0; j < 4; j++) a[j] = f_add(a[j], a[j + 4]);
                a[0] = f_add(a[0], a[2]);
                a[1] = f_add(a[1], a[3]);
                float v = f_add(a[0], a[1]);
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) v = f_add(v, __shfl_down_sync(0xffffffffu, v, o));
                if (lane == 0) {
                    if (N == 1) v = (q == 0) ? wgt[0] : f_mul(wgt[0], space[q - 1]);  // no reduction pass at all
                    partials[(size_t)q * kMaxTreeBlocks + b] = v;
                }
            }
        }
        __syncthreads();
        // level 2: one warp per quantity over the NB block results
        for (int q = warp; q < (A.trial_only ? 1 : Q); q += kMsWarps) {
            const float* p = partials + (size_t)q * kMaxTreeBlocks;
            float v;
            if (NB == 1)
                v = p[0];
            else
                v = tree_sum_512([&](int i) { return p[i]; }, NB, lane);
            if (lane == 0) sums[q] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            const float wsum = sums[0];
            wsum_last = wsum;
            if (!A.trial_only) {
                // host part of the reference iteration (meanshift.cu:112-133)
                float mean_new[kMeanshiftMaxDims];
                for (int d = 0; d < dims; d++) mean_new[d] = f_div(sums[d + 1], wsum);
                confidence = f_div(wsum, (float)N);
                used = iter + 1;
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
    if (threadIdx.x == 0) {
        for (int d = 0; d < dims; d++) out->mean[d] = io_mean[d];
        out->confidence = confidence;
        out->weight_sum = wsum_last;
        out->used_iters = used;
        out->n = N;
    }
}

// ------------------------------------------------------------------------------------------------
// robust Gaussian: E-step + all weighted sums in one launch (reference: fit_robust_gaussian.cu:56-97,211-242)
// ------------------------------------------------------------------------------------------------
struct RobustArgs {
    float mean[kRobustMaxDims];
    float cinv[21];  // lower-triangular packed inverse covariance
    float trunc_sigma, scale;
    int N, dims;
};

constexpr int kRgThreads = 1024;
constexpr int kRgWarps = kRgThreads / 32;

__global__ void __launch_bounds__(kRgThreads)
    k_robust_estep(const float* __restrict__ space, const RobustArgs A, float* scratch, float* partials,
                   float* sums_out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N = A.N, dims = A.dims;
    const int cdims = (dims * dims + dims) / 2;
    const int Q = 1 + dims + cdims;
    float* W = scratch;                      // [N]
    float* WS = scratch + N;                 // [N][dims]
    float* WC = scratch + (size_t)N * (1 + dims);  // [N][cdims]

    for (int idx = threadIdx.x; idx < N; idx += kRgThreads) {
        float diff[kRobustMaxDims], x[kRobustMaxDims];
        for (int d = 0; d < dims; d++) {
            x[d] = f_mul(space[(size_t)idx * dims + d], A.scale);
            diff[d] = f_sub(x[d], A.mean[d]);
        }
        float z = 0.f;
        for (int d1 = 0; d1 < dims; d1++) {
            float tmp = 0.f;
            for (int d2 = 0; d2 < dims; d2++) {
                const float ci = (d1 >= d2) ? A.cinv[(d1 * d1 + d1) / 2 + d2] : A.cinv[(d2 * d2 + d2) / 2 + d1];
                tmp = f_add(tmp, f_mul(ci, diff[d2]));
            }
            z = f_fma(tmp, diff[d1], z);
        }
        z = __fsqrt_rn(z);
        const float weight = z < A.trunc_sigma ? 1.f : 0.f;  // hard truncation
        W[idx] = weight;
        for (int d = 0; d < dims; d++) WS[(size_t)idx * dims + d] = f_mul(weight, x[d]);
        for (int d1 = 0; d1 < dims; d1++)
            for (int d2 = 0; d2 <= d1; d2++)
                WC[(size_t)idx * cdims + (d1 * d1 + d1) / 2 + d2] = f_mul(f_mul(weight, diff[d1]), diff[d2]);
    }
    __syncthreads();

    const int NB = (N + 511) / 512;
    for (int task = warp; task < Q * NB; task += kRgWarps) {
        const int q = task / NB, b = task % NB;
        const float* src;
        int stride;
        if (q == 0)
            src = W, stride = 1;
        else if (q <= dims)
            src = WS + (q - 1), stride = dims;
        else
            src = WC + (q - 1 - dims), stride = cdims;
        const int base = b * 512;
        const int count = min(512, N - base);
        float v = tree_sum_512([&](int i) { return src[(size_t)(base + i) * stride]; }, count, lane);
        if (lane == 0) {
            if (N == 1) v = src[0];
            partials[(size_t)q * kMaxTreeBlocks + b] = v;
        }
    }
    __syncthreads();
    for (int q = warp; q < Q; q += kRgWarps) {
        const float* p = partials + (size_t)q * kMaxTreeBlocks;
        float v = (NB == 1) ? p[0] : tree_sum_512([&](int i) { return p[i]; }, NB, lane);
        if (lane == 0) sums_out[q] = v;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
int PoseMode::init() {
    if (stream) return 0;
    VB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    VB_CUDA(cudaMalloc((void**)&d_result, sizeof(MeanshiftResult)));
    VB_CUDA(cudaMallocHost((void**)&h_result, sizeof(MeanshiftResult)));
    VB_CUDA(cudaMalloc((void**)&d_partials, (size_t)64 * kMaxTreeBlocks * sizeof(float)));
    VB_CUDA(cudaMalloc((void**)&d_rg_sums, 64 * sizeof(float)));
    VB_CUDA(cudaMallocHost((void**)&h_rg_sums, 64 * sizeof(float)));
    return 0;
}

int PoseMode::meanshift(const float* d_space, const float* h_space_for_init, const int* d_n, int n_host, int dims,
                        float kernel_var, float* h_io_mean, float* h_o_confidence, int* used_iters,
                        bool use_external_init_mean, float epsilon, int max_iters, int max_init_trials,
                        float good_init_confidence) {
    (void)h_space_for_init;
    if (int e = init()) return e;
    if (dims > kMeanshiftMaxDims) return (int)cudaErrorInvalidValue;
    MeanshiftArgs A;
    memset(&A, 0, sizeof(A));
    for (int d = 0; d < dims; d++) A.io_mean[d] = h_io_mean[d];
    A.n_host = n_host, A.dims = dims, A.kernel_var = kernel_var, A.epsilon = epsilon, A.max_iters = max_iters;
    A.center_idx = -1, A.trial_only = 0;

    int N = n_host;
    if (!use_external_init_mean) {
        // start from the best of <= max_init_trials samples drawn with the host libc rand() (meanshift.cu:73-97;
        // the unseeded process-wide stream is part of the reference behaviour)
        if (d_n) {
            VB_CUDA(cudaMemcpyAsync(&h_result->n, d_n, sizeof(int), cudaMemcpyDeviceToHost, stream));
            VB_CUDA(cudaStreamSynchronize(stream));
            N = h_result->n;
        }
        if (N > 512 * 512) return (int)cudaErrorInvalidValue;
        float best_conf = 0;
        int best_idx = -1;
        for (int trial = 0; trial < max_init_trials; trial++) {
            const int idx_rand = (rand() % N);
            MeanshiftArgs T = A;
            T.center_idx = idx_rand, T.trial_only = 1, T.n_host = N;
            k_meanshift<<<1, kMsThreads, 0, stream>>>(d_space, nullptr, T, d_partials, d_result);
            VB_RETURN_IF_CUDA_ERROR();
            VB_CUDA(cudaMemcpyAsync(h_result, d_result, sizeof(MeanshiftResult), cudaMemcpyDeviceToHost, stream));
            VB_CUDA(cudaStreamSynchronize(stream));
            if (h_result->weight_sum > best_conf) {
                best_conf = h_result->weight_sum;
                best_idx = idx_rand;
            }
            if (best_conf > good_init_confidence * N) break;
        }
        A.center_idx = best_idx < 0 ? 0 : best_idx;
        A.n_host = N;
        d_n = nullptr;
    } else if (!d_n && N > 512 * 512) {
        return (int)cudaErrorInvalidValue;
    }

    if (used_iters) *used_iters = 0;
    k_meanshift<<<1, kMsThreads, 0, stream>>>(d_space, d_n, A, d_partials, d_result);
    VB_RETURN_IF_CUDA_ERROR();
    VB_CUDA(cudaMemcpyAsync(h_result, d_result, sizeof(MeanshiftResult), cudaMemcpyDeviceToHost, stream));
    VB_CUDA(cudaStreamSynchronize(stream));
    if (h_result->used_iters > 0) {
        if (h_o_confidence) *h_o_confidence = h_result->confidence;
        if (used_iters) *used_iters = h_result->used_iters;
        for (int d = 0; d < dims; d++) h_io_mean[d] = h_result->mean[d];
    }
    return 0;
}

int PoseMode::fit_robust_gaussian(const float* d_space, int N, int dims, float scale, float* h_io_mean,
                                  float* h_io_covar, float trunc_sigma, float covar_reg_lambda, float* h_o_density,
                                  int* used_iters, float epsilon, int max_iters) {
    if (int e = init()) return e;
    if (dims > kRobustMaxDims) throw;  // reference: fit_robust_gaussian.cu:107-108
    if (N > 512 * 512) return (int)cudaErrorInvalidValue;
    const int cdims = (dims * dims + dims) / 2;
    const int Q = 1 + dims + cdims;
    const size_t need = (size_t)N * Q;
    if (need > rg_capacity) {
        if (d_rg_scratch) cudaFree(d_rg_scratch);
        VB_CUDA(cudaMalloc((void**)&d_rg_scratch, need * sizeof(float)));
        rg_capacity = need;
    }

    float ht_weight = 0;
    float ht_mean[kRobustMaxDims];
    float ht_covar[21], ht_covar_inv[21];
    double covar_full[36], covar_inv_full[36];
    for (int d = 0; d < dims; d++) ht_mean[d] = h_io_mean[d];
    for (int d1 = 0; d1 < dims; d1++)
        for (int d2 = 0; d2 <= d1; d2++) ht_covar[(d1 * d1 + d1) / 2 + d2] = h_io_covar[d1 * dims + d2];
    if (used_iters) *used_iters = 0;

    int iter;
    bool reliable = true;
    for (iter = 0; iter < max_iters; iter++) {
        // half -> full (double), regularise from the 2nd iteration on, invert (fit_robust_gaussian.cu:166-201)
        for (int d1 = 0; d1 < dims; d1++)
            for (int d2 = 0; d2 <= d1; d2++) {
                covar_full[d1 * dims + d2] = (double)ht_covar[(d1 * d1 + d1) / 2 + d2];
                if (d1 != d2) covar_full[d2 * dims + d1] = covar_full[d1 * dims + d2];
            }
        if (iter > 0 && covar_reg_lambda > 0) linalg::shrink_to_scaled_identity6(covar_full, covar_reg_lambda, dims);
        const double det = linalg::inverse6(covar_full, covar_inv_full, dims);
        if (det <= 0) {
            reliable = false;
            break;
        }
        for (int d1 = 0; d1 < dims; d1++)
            for (int d2 = 0; d2 <= d1; d2++) {
                ht_covar[(d1 * d1 + d1) / 2 + d2] = (float)covar_full[d1 * dims + d2];
                ht_covar_inv[(d1 * d1 + d1) / 2 + d2] = (float)covar_inv_full[d1 * dims + d2];
            }

        RobustArgs A;
        memset(&A, 0, sizeof(A));
        for (int d = 0; d < dims; d++) A.mean[d] = ht_mean[d];
        for (int k = 0; k < cdims; k++) A.cinv[k] = ht_covar_inv[k];
        A.trunc_sigma = trunc_sigma, A.scale = scale, A.N = N, A.dims = dims;

        const float prev_density = ht_weight / N;
        k_robust_estep<<<1, kRgThreads, 0, stream>>>(d_space, A, d_rg_scratch, d_partials, d_rg_sums);
        VB_RETURN_IF_CUDA_ERROR();
        VB_CUDA(cudaMemcpyAsync(h_rg_sums, d_rg_sums, Q * sizeof(float), cudaMemcpyDeviceToHost, stream));
        VB_CUDA(cudaStreamSynchronize(stream));

        ht_weight = h_rg_sums[0];
        if (!std::isfinite(ht_weight)) {
            reliable = false;
            break;
        }
        const float density_change = std::fabs(ht_weight / N - prev_density);
        if (density_change < epsilon) {
            reliable = true;
            break;
        }
        // not converged: adopt the new moments (Q15: on convergence the moments USED in the last E-step stay)
        for (int d = 0; d < dims; d++) ht_mean[d] = h_rg_sums[1 + d] / ht_weight;
        for (int k = 0; k < cdims; k++) ht_covar[k] = h_rg_sums[1 + dims + k] / ht_weight;
    }

    if (reliable) {
        if (h_o_density) *h_o_density = ht_weight / N;
        if (used_iters) *used_iters = iter;
        for (int d1 = 0; d1 < dims; d1++)
            for (int d2 = 0; d2 <= d1; d2++) {
                h_io_covar[d1 * dims + d2] = ht_covar[(d1 * d1 + d1) / 2 + d2];
                h_io_covar[d2 * dims + d1] = h_io_covar[d1 * dims + d2];
            }
        for (int d = 0; d < dims; d++) h_io_mean[d] = ht_mean[d];
    }
    return reliable ? 0 : 1;  // cudaSuccess / !cudaSuccess (fit_robust_gaussian.cu:281-284)
}

PoseMode& global_pose_mode() {
    static PoseMode inst;
    return inst;
}

}  // namespace vb
