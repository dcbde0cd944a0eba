// Library-level drop-in boundary: the reference's gpu_kernels.h entry points on top of the sm_100a kernels.
//
// Each function keeps the host-pointer / NULL-means-cached / return-code contract of the reference
// (gpu-kernels/gpu_kernels.h:11-74) and is a thin wrapper: stage host data into the
// device-resident state objects, enqueue the kernels, download what the caller asked for.
#include "../../include/gpu_kernels.h"
#include "../../include/voldor_b200.h"
#include "context.h"
#include "host_math.h"
#include <mutex>

// Every entry point works on the calling host thread's execution context (context.h; context 0 unless the thread
// selected another one) and holds that context's mutex for the duration of the call.
#define VB_ENTER_CONTEXT()                    \
    vb::enter_device();                       \
    vb::Context& cx = vb::current_context(); \
    std::lock_guard<std::recursive_mutex> lock(cx.mutex)

// ---------------------------------------------------------------------------------------------------
int optimize_depth_gpu(float* h_flows[], float* h_rigidnesses[], float* h_o_rigidnesses[], float* h_depth_priors[],
                       float* h_depth_prior_pconfs[], float* h_depth_prior_confs[], float* h_o_depth_prior_confs[],
                       float* h_depth, float* h_o_depth, float* h_K, float* h_Rs[], float* h_ts[], float* h_dp_Rs[],
                       float* h_dp_ts[], float abs_resize_factor, int N, int N_dp, int w, int h, float basefocal,
                       int n_rand_samples, int global_prop_step, int local_prop_width, float lambda, float omega,
                       float disp_delta, float delta, bool fb_smooth, float s0_ems_prob, float no_change_prob,
                       float range_factor, bool update_rigidness_only) {
    VB_ENTER_CONTEXT();
    if (N > vb::kMaxFrames || N_dp > vb::kMaxPriorFrames) return (int)cudaErrorInvalidValue;
    vb::DepthEM& E = cx.E;
    E.shared_flows = nullptr;
    E.overlap_smoothing = false;  // ABI calls hand the raw maps back to the host after every step
    if (int e = E.ensure(w, h, N, N_dp)) return e;
    cudaStream_t s = E.stream;

    if (h_K) E.set_K(h_K);
    if (h_depth) VB_CUDA(E.depth.upload_layer(h_depth, 0, s));
    if (N > 0) {
        if (h_Rs)
            for (int f = 0; f < N; f++) memcpy(E.cam.R[f], h_Rs[f], 9 * sizeof(float));
        if (h_ts)
            for (int f = 0; f < N; f++) memcpy(E.cam.t[f], h_ts[f], 3 * sizeof(float));
        if (h_flows)
            for (int f = 0; f < N; f++) VB_CUDA(E.flows.upload_layer((const float2*)h_flows[f], f, s));
        if (h_rigidnesses)
            for (int f = 0; f < N; f++) VB_CUDA(E.rig.upload_layer(h_rigidnesses[f], f, s));
    }
    if (N_dp > 0) {
        if (h_dp_Rs)
            for (int f = 0; f < N_dp; f++) memcpy(E.pcam.R[f], h_dp_Rs[f], 9 * sizeof(float));
        if (h_dp_ts)
            for (int f = 0; f < N_dp; f++) memcpy(E.pcam.t[f], h_dp_ts[f], 3 * sizeof(float));
        if (h_depth_priors)
            for (int f = 0; f < N_dp; f++) VB_CUDA(E.dp.upload_layer(h_depth_priors[f], f, s));
        if (h_depth_prior_pconfs)
            for (int f = 0; f < N_dp; f++) VB_CUDA(E.dp_pconf.upload_layer(h_depth_prior_pconfs[f], f, s));
        if (h_depth_prior_confs)
            for (int f = 0; f < N_dp; f++) VB_CUDA(E.dp_conf.upload_layer(h_depth_prior_confs[f], f, s));
    }

    vb::DepthHyper hp;
    hp.abs_resize_factor = abs_resize_factor, hp.basefocal = basefocal;
    hp.n_rand_samples = n_rand_samples, hp.global_prop_step = global_prop_step;
    hp.local_prop_width = local_prop_width;
    hp.lambda = lambda, hp.omega = omega, hp.disp_delta = disp_delta, hp.delta = delta;
    hp.fb_smooth = fb_smooth, hp.s0_ems_prob = s0_ems_prob, hp.no_change_prob = no_change_prob;
    hp.range_factor = range_factor;
    if (int e = E.run(N, N_dp, hp, update_rigidness_only)) return e;

    if (h_o_depth) VB_CUDA(E.depth.download_layer(h_o_depth, 0, s));
    if (h_o_rigidnesses)
        for (int f = 0; f < N; f++) VB_CUDA(E.rig.download_layer(h_o_rigidnesses[f], f, s));
    if (h_o_depth_prior_confs)
        for (int f = 0; f < N_dp; f++) VB_CUDA(E.dp_conf.download_layer(h_o_depth_prior_confs[f], f, s));
    VB_CUDA(cudaStreamSynchronize(s));
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
int collect_p3p_instances(float* h_flows[], float* h_rigidnesses[], float* h_depth, float* h_K, float* h_Rs[],
                          float* h_ts[], float* h_o_p2_map, float* h_o_p3_map, int N, int w, int h, int active_idx,
                          float rigidness_thresh, float rigidness_sum_thresh, float sample_min_depth,
                          float sample_max_depth, int max_trace_on_flow) {
    VB_ENTER_CONTEXT();
    if (N > vb::kMaxFrames) return (int)cudaErrorInvalidValue;
    vb::Collector& C = cx.C;
    if (int e = C.ensure(w, h, N)) return e;
    cudaStream_t s = C.stream;
    C.flows_own.ensure(w, h, N, true);
    C.rig_own.ensure(w, h, N, true);
    C.depth_own.ensure(w, h, 1, false);
    C.use_own_views();

    if (h_K) vb::fill_K(C.cam, h_K);
    if (h_Rs)
        for (int f = 0; f < N; f++) memcpy(C.cam.R[f], h_Rs[f], 9 * sizeof(float));
    if (h_ts)
        for (int f = 0; f < N; f++) memcpy(C.cam.t[f], h_ts[f], 3 * sizeof(float));
    if (h_flows)
        for (int f = 0; f < N; f++) VB_CUDA(C.flows_own.upload_layer((const float2*)h_flows[f], f, s));
    if (h_rigidnesses)
        for (int f = 0; f < N; f++) VB_CUDA(C.rig_own.upload_layer(h_rigidnesses[f], f, s));
    if (h_depth) VB_CUDA(C.depth_own.upload_layer(h_depth, 0, s));

    vb::CollectParams P;
    P.active_idx = active_idx, P.rigidness_thresh = rigidness_thresh, P.rigidness_sum_thresh = rigidness_sum_thresh;
    P.sample_min_depth = sample_min_depth, P.sample_max_depth = sample_max_depth;
    P.max_trace_on_flow = max_trace_on_flow;
    if (int e = C.collect(N, P, false)) return e;

    const size_t npx = (size_t)w * h;
    if (h_o_p2_map)
        VB_CUDA(cudaMemcpyAsync(h_o_p2_map, C.p2_map, npx * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    if (h_o_p3_map)
        VB_CUDA(cudaMemcpyAsync(h_o_p3_map, C.p3_map, npx * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaStreamSynchronize(s));
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
static int solve_batch_host(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K, int N_pts,
                            int N_poses, bool ap3p) {
    VB_ENTER_CONTEXT();
    vb::DevBuf &p2 = cx.abi.p2, &p3 = cx.abi.p3, &rv = cx.abi.rv, &tv = cx.abi.tv;
    float* K4 = cx.abi.K4;
    if (!cx.abi.p3p_stream) VB_CUDA(cudaStreamCreateWithFlags(&cx.abi.p3p_stream, cudaStreamNonBlocking));
    cudaStream_t s = cx.abi.p3p_stream;
    if (h_K) K4[0] = h_K[0], K4[1] = h_K[4], K4[2] = h_K[2], K4[3] = h_K[5];
    if (p2.ensure(((size_t)N_pts + 1) * 2) || p3.ensure(((size_t)N_pts + 1) * 3) || rv.ensure((size_t)N_poses * 3) ||
        tv.ensure((size_t)N_poses * 3))
        return (int)cudaErrorMemoryAllocation;
    // one padded element: an index can equal N_pts; the reference reads past the end there
    VB_CUDA(cudaMemsetAsync(p2.ptr + (size_t)N_pts * 2, 0, 2 * sizeof(float), s));
    VB_CUDA(cudaMemsetAsync(p3.ptr + (size_t)N_pts * 3, 0, 3 * sizeof(float), s));
    VB_CUDA(cudaMemcpyAsync(p2.ptr, h_p2s, (size_t)N_pts * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
    VB_CUDA(cudaMemcpyAsync(p3.ptr, h_p3s, (size_t)N_pts * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (int e = vb::solve_batch_p3p_device(p3.ptr, p2.ptr, nullptr, N_pts, K4[0], K4[1], K4[2], K4[3], rv.ptr, tv.ptr,
                                           N_poses, ap3p, s))
        return e;
    VB_CUDA(cudaMemcpyAsync(h_o_rvecs, rv.ptr, (size_t)N_poses * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaMemcpyAsync(h_o_tvecs, tv.ptr, (size_t)N_poses * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    VB_CUDA(cudaStreamSynchronize(s));
    VB_RETURN_IF_CUDA_ERROR();
    return 0;
}

int solve_batch_p3p_ap3p_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K, int N_pts,
                             int N_poses) {
    return solve_batch_host(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses, true);
}
int solve_batch_p3p_lambdatwist_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K,
                                    int N_pts, int N_poses) {
    return solve_batch_host(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses, false);
}

// ---------------------------------------------------------------------------------------------------
int meanshift_gpu(float* h_space, float kernel_var, float* h_io_mean, float* h_o_confidence, int* used_iters,
                  bool use_external_init_mean, int N, int dims, float epsilon, int max_iters, int max_init_trials,
                  float good_init_confidence) {
    VB_ENTER_CONTEXT();
    vb::DevBuf& space = cx.abi.ms_space;
    vb::PoseMode& M = cx.M;
    if (int e = M.init()) return e;
    if (space.ensure((size_t)N * dims)) return (int)cudaErrorMemoryAllocation;
    VB_CUDA(cudaMemcpyAsync(space.ptr, h_space, (size_t)N * dims * sizeof(float), cudaMemcpyHostToDevice, M.stream));
    return M.meanshift(space.ptr, h_space, nullptr, N, dims, kernel_var, h_io_mean, h_o_confidence, used_iters,
                       use_external_init_mean, epsilon, max_iters, max_init_trials, good_init_confidence);
}

int fit_robust_gaussian(float* h_space, float* h_io_mean, float* h_io_covar, float trunc_sigma,
                        float covar_reg_lambda, float* h_o_density, int* used_iters, int N, int dims, float epsilon,
                        int max_iters) {
    VB_ENTER_CONTEXT();
    vb::DevBuf& space = cx.abi.rg_space;
    vb::PoseMode& M = cx.M;
    if (int e = M.init()) return e;
    if (space.ensure((size_t)N * dims)) return (int)cudaErrorMemoryAllocation;
    VB_CUDA(cudaMemcpyAsync(space.ptr, h_space, (size_t)N * dims * sizeof(float), cudaMemcpyHostToDevice, M.stream));
    return M.fit_robust_gaussian(space.ptr, N, dims, 1.0f, h_io_mean, h_io_covar, trunc_sigma, covar_reg_lambda,
                                 h_o_density, used_iters, epsilon, max_iters);
}

// ---------------------------------------------------------------------------------------------------
// C linkage aliases
// ---------------------------------------------------------------------------------------------------
extern "C" {

DLL_EXPORT int vb_meanshift_gpu(float* h_space, float kernel_var, float* h_io_mean, float* h_o_confidence,
                                int* used_iters, int use_external_init_mean, int N, int dims, float epsilon,
                                int max_iters, int max_init_trials, float good_init_confidence) {
    return meanshift_gpu(h_space, kernel_var, h_io_mean, h_o_confidence, used_iters, use_external_init_mean != 0, N,
                         dims, epsilon, max_iters, max_init_trials, good_init_confidence);
}
DLL_EXPORT int vb_fit_robust_gaussian(float* h_space, float* h_io_mean, float* h_io_covar, float trunc_sigma,
                                      float covar_reg_lambda, float* h_o_density, int* used_iters, int N, int dims,
                                      float epsilon, int max_iters) {
    return fit_robust_gaussian(h_space, h_io_mean, h_io_covar, trunc_sigma, covar_reg_lambda, h_o_density, used_iters,
                               N, dims, epsilon, max_iters);
}
DLL_EXPORT int vb_collect_p3p_instances(float** h_flows, float** h_rigidnesses, float* h_depth, float* h_K,
                                        float** h_Rs, float** h_ts, float* h_o_p2_map, float* h_o_p3_map, int N,
                                        int w, int h, int active_idx, float rigidness_thresh,
                                        float rigidness_sum_thresh, float sample_min_depth, float sample_max_depth,
                                        int max_trace_on_flow) {
    return collect_p3p_instances(h_flows, h_rigidnesses, h_depth, h_K, h_Rs, h_ts, h_o_p2_map, h_o_p3_map, N, w, h,
                                 active_idx, rigidness_thresh, rigidness_sum_thresh, sample_min_depth,
                                 sample_max_depth, max_trace_on_flow);
}
DLL_EXPORT int vb_solve_batch_p3p_ap3p_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs,
                                           float* h_K, int N_pts, int N_poses) {
    return solve_batch_p3p_ap3p_gpu(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses);
}
DLL_EXPORT int vb_solve_batch_p3p_lambdatwist_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs,
                                                  float* h_K, int N_pts, int N_poses) {
    return solve_batch_p3p_lambdatwist_gpu(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses);
}
DLL_EXPORT int vb_optimize_depth_gpu(float** h_flows, float** h_rigidnesses, float** h_o_rigidnesses,
                                     float** h_depth_priors, float** h_depth_prior_pconfs,
                                     float** h_depth_prior_confs, float** h_o_depth_prior_confs, float* h_depth,
                                     float* h_o_depth, float* h_K, float** h_Rs, float** h_ts, float** h_dp_Rs,
                                     float** h_dp_ts, float abs_resize_factor, int N, int N_dp, int w, int h,
                                     float basefocal, int n_rand_samples, int global_prop_step, int local_prop_width,
                                     float lambda, float omega, float disp_delta, float delta, int fb_smooth,
                                     float s0_ems_prob, float no_change_prob, float range_factor,
                                     int update_rigidness_only) {
    return optimize_depth_gpu(h_flows, h_rigidnesses, h_o_rigidnesses, h_depth_priors, h_depth_prior_pconfs,
                              h_depth_prior_confs, h_o_depth_prior_confs, h_depth, h_o_depth, h_K, h_Rs, h_ts, h_dp_Rs,
                              h_dp_ts, abs_resize_factor, N, N_dp, w, h, basefocal, n_rand_samples, global_prop_step,
                              local_prop_width, lambda, omega, disp_delta, delta, fb_smooth != 0, s0_ems_prob,
                              no_change_prob, range_factor, update_rigidness_only != 0);
}

DLL_EXPORT int vb_set_device(int device) { return vb::set_library_device(device); }
DLL_EXPORT void vb_profile_enable(int on) {
    vb::KernelProfile& p = vb::current_context().prof;
    p.enabled = on != 0;
    p.search_ms = 0, p.search_launches = 0;
    p.estep_ms = p.smooth_ms = p.local_ms = 0, p.estep_launches = p.smooth_runs = p.local_runs = 0;
    p.meanshift_runs = p.meanshift_iters = p.meanshift_trials = p.robust_runs = p.robust_iters = 0;
}
DLL_EXPORT void vb_profile_counters(long long* out5) {
    vb::KernelProfile& p = vb::current_context().prof;
    out5[0] = p.meanshift_runs, out5[1] = p.meanshift_iters, out5[2] = p.meanshift_trials;
    out5[3] = p.robust_runs, out5[4] = p.robust_iters;
}
DLL_EXPORT void vb_profile_get(double* search_ms, long long* search_launches) {
    vb::KernelProfile& p = vb::current_context().prof;
    if (search_ms) *search_ms = p.search_ms;
    if (search_launches) *search_launches = p.search_launches;
}
DLL_EXPORT void vb_profile_get_more(double* ms3, long long* counts3) {
    vb::KernelProfile& p = vb::current_context().prof;
    ms3[0] = p.estep_ms, ms3[1] = p.smooth_ms, ms3[2] = p.local_ms;
    counts3[0] = p.estep_launches, counts3[1] = p.smooth_runs, counts3[2] = p.local_runs;
}
DLL_EXPORT int vb_debug_pose_mode_phases(long long* out24) {
    vb::PoseMode& M = vb::current_context().M;
    if (!M.d_phase_cycles) return 1;
    return (int)cudaMemcpy(out24, M.d_phase_cycles, 24 * sizeof(long long), cudaMemcpyDeviceToHost);
}
// rvec -> R through the device code path and through the host code path of the same source (csrc/host_math.h)
namespace {
__global__ void k_debug_rvec_to_matrix(const float* rvecs, int n, float* R) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) vb::hm::rvec_to_matrix(rvecs + 3 * i, R + 9 * i);
}
}  // namespace
DLL_EXPORT int vb_debug_rvec_to_matrix(const float* rvecs, int n, float* R_device, float* R_host) {
    float *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc((void**)&d_in, (size_t)n * 3 * sizeof(float)) != cudaSuccess) return 1;
    if (cudaMalloc((void**)&d_out, (size_t)n * 9 * sizeof(float)) != cudaSuccess) return 1;
    cudaMemcpy(d_in, rvecs, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice);
    k_debug_rvec_to_matrix<<<(n + 127) / 128, 128>>>(d_in, n, d_out);
    const cudaError_t e = cudaMemcpy(R_device, d_out, (size_t)n * 9 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_in), cudaFree(d_out);
    for (int i = 0; i < n; i++) vb::hm::rvec_to_matrix(rvecs + 3 * i, R_host + 9 * i);
    return (int)e;
}
DLL_EXPORT int vb_debug_rand_speculate(int draw, int keep) {
    vb::Context& cx = vb::current_context();
    std::lock_guard<std::recursive_mutex> lock(cx.mutex);
    vb::RandStream& r = *cx.rnd;
    if (!r.snapshot()) return 1;
    for (int i = 0; i < draw; i++) (void)r.next();
    r.rewind();
    for (int i = 0; i < keep; i++) (void)r.next();
    return 0;
}
DLL_EXPORT const char* vb_version(void) { return "voldor_b200 0.1 sm_100a"; }

}  // extern "C"
