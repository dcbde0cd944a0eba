// Mode seeking over the pose-hypothesis pool: Gaussian mean-shift and truncated robust Gaussian fit.
//
// Replaces reference gpu-kernels/meanshift.cu:34-150 (host loop of <=100 iterations, each 1 kernel +
// 2 multi-pass reductions + 2 blocking D2H + 1 cudaMemcpyToSymbol, preceded by <=20 start-sample trials of the
// same shape) and reference gpu-kernels/fit_robust_gaussian.cu:101-286 + aux_funs.cpp:101-141 (host loop with a
// 6x6 FP64 determinant/inverse/shrinkage per iteration) with ONE launch each of an 8-CTA thread-block cluster
// that runs the whole fixed-point loop on the device (pose_mode.cu).
#pragma once
#include "common.cuh"

namespace vb {

constexpr int kMeanshiftMaxDims = 16;  // reference: meanshift.cu:5
constexpr int kRobustMaxDims = 6;      // reference: fit_robust_gaussian.cu:6

struct MeanshiftResult {
    float mean[kMeanshiftMaxDims];
    float confidence;
    float weight_sum;  // last sum of kernel weights (trial mode: the trial's weight sum)
    int used_iters;
    int n;             // pool size seen by the kernel
    int aux_count;     // echo of the caller's device counter (number of P3P instances in the window pipeline)
    int trials_used;   // start-sample trials the reference's selection loop would have run (fused trial mode)
    // pose tail (window pipeline, see PoseTail): camera pose derived from the mode on the device
    int ok;
    float R[9], t[3];
};

// Optional tail of the mean-shift launch in the window pipeline: turn the mode into the camera's rotation matrix and
// translation ON THE DEVICE (reference voldor/geometry.cpp:191-262: un-scale the rotation vector, finite check,
// Rodrigues) and store it in the device-resident camera block, so the next camera's kernels of the same EM
// iteration can be enqueued without waiting for the host.
struct PoseTail {
    CamBlock* d_cams;
    int cam_index;
    float inv_rvec_scale;
};

struct RandStream;
struct KernelProfile;

struct PoseMode {
    RandStream* rnd = nullptr;      // start-sample stream of the owning context (libc_rand.h); never null when used
    KernelProfile* prof = nullptr;  // the owning context's counters
    cudaStream_t stream = nullptr;
    MeanshiftResult* d_result = nullptr;  // device, kMaxFrames slots (slot 0 is the synchronous API's)
    MeanshiftResult* h_result = nullptr;  // pinned host mirror
    float* d_partials = nullptr;
    // robust fit scratch
    float* d_rg_scratch = nullptr;
    size_t rg_capacity = 0;
    float* d_rg_sums = nullptr;  // RobustResult on the device
    float* h_rg_sums = nullptr;  // pinned mirror
    long long* d_phase_cycles = nullptr;  // robust-fit phase clocks, allocated when VB_POSE_MODE_PHASES is set

    int init();
    // Mean-shift on d_space[N][dims]; N is read from d_n when non-null.  Blocks until the result is on the host
    // (needed by the caller's control flow).  Semantics of every argument as in the reference ABI.
    int meanshift(const float* d_space, const float* h_space_for_init, const int* d_n, int n_host, int dims,
                  float kernel_var, float* h_io_mean, float* h_o_confidence, int* used_iters,
                  bool use_external_init_mean, float epsilon, int max_iters, int max_init_trials,
                  float good_init_confidence, int n_capacity = 0);
    // Successive-pose mean-shift fused with the finite filter of the raw hypotheses: builds the pool
    // (hypothesis order, rvec * rvec_scale) into d_pool / d_used and iterates on it in the same launch.
    int meanshift_from_hypotheses(const float* d_rvecs, const float* d_tvecs, int n_poses, float rvec_scale,
                                  float* d_pool, int* d_used, int dims, float kernel_var, float* h_io_mean,
                                  float* h_o_confidence, int* used_iters, float epsilon, int max_iters,
                                  const int* d_aux_count = nullptr);
    // Asynchronous variant for the window pipeline: enqueue only, result (incl. the pose tail) goes to slot `slot`;
    // fetch_results(n) copies slots [0,n) back and synchronises once.
    int enqueue_from_hypotheses(int slot, const float* d_rvecs, const float* d_tvecs, int n_poses, float rvec_scale,
                                float* d_pool, int* d_used, int dims, float kernel_var, const float* h_init_mean,
                                float epsilon, int max_iters, const int* d_aux_count, const PoseTail& tail);
    int fetch_results(int n_slots);
    // Robust Gaussian fit on x = scale * d_space (scale folds the caller's pose scaling).
    int fit_robust_gaussian(const float* d_space, int N, int dims, float scale, float* h_io_mean,
                            float* h_io_covar, float trunc_sigma, float covar_reg_lambda, float* h_o_density,
                            int* used_iters, float epsilon, int max_iters);
};

}  // namespace vb
