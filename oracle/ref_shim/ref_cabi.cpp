// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// extern "C" trampolines (prefix ref_) onto the reference library's C++-linkage entry points
// (reference: gpu-kernels/gpu_kernels.h:11-74), so that the test harness can dlopen/ctypes the
// reference build next to ours inside one process.  gpu_kernels.h is taken from the reference tree by
// -I at build time (see oracle/Makefile); nothing from the reference is copied into this repository.
#include "gpu_kernels.h"

extern "C" {

int ref_meanshift_gpu(float* h_space, float kernel_var, float* h_io_mean, float* h_o_confidence, int* used_iters,
                      int use_external_init_mean, int N, int dims, float epsilon, int max_iters, int max_init_trials,
                      float good_init_confidence) {
    return meanshift_gpu(h_space, kernel_var, h_io_mean, h_o_confidence, used_iters, use_external_init_mean != 0, N,
                         dims, epsilon, max_iters, max_init_trials, good_init_confidence);
}

int ref_fit_robust_gaussian(float* h_space, float* h_io_mean, float* h_io_covar, float trunc_sigma,
                            float covar_reg_lambda, float* h_o_density, int* used_iters, int N, int dims,
                            float epsilon, int max_iters) {
    return fit_robust_gaussian(h_space, h_io_mean, h_io_covar, trunc_sigma, covar_reg_lambda, h_o_density, used_iters,
                               N, dims, epsilon, max_iters);
}

int ref_collect_p3p_instances(float** h_flows, float** h_rigidnesses, float* h_depth, float* h_K, float** h_Rs,
                              float** h_ts, float* h_o_p2_map, float* h_o_p3_map, int N, int w, int h, int active_idx,
                              float rigidness_thresh, float rigidness_sum_thresh, float sample_min_depth,
                              float sample_max_depth, int max_trace_on_flow) {
    return collect_p3p_instances(h_flows, h_rigidnesses, h_depth, h_K, h_Rs, h_ts, h_o_p2_map, h_o_p3_map, N, w, h,
                                 active_idx, rigidness_thresh, rigidness_sum_thresh, sample_min_depth,
                                 sample_max_depth, max_trace_on_flow);
}

int ref_solve_batch_p3p_ap3p_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K,
                                 int N_pts, int N_poses) {
    return solve_batch_p3p_ap3p_gpu(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses);
}

int ref_solve_batch_p3p_lambdatwist_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K,
                                        int N_pts, int N_poses) {
    return solve_batch_p3p_lambdatwist_gpu(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses);
}

int ref_optimize_depth_gpu(float** h_flows, float** h_rigidnesses, float** h_o_rigidnesses, float** h_depth_priors,
                           float** h_depth_prior_pconfs, float** h_depth_prior_confs, float** h_o_depth_prior_confs,
                           float* h_depth, float* h_o_depth, float* h_K, float** h_Rs, float** h_ts, float** h_dp_Rs,
                           float** h_dp_ts, float abs_resize_factor, int N, int N_dp, int w, int h, float basefocal,
                           int n_rand_samples, int global_prop_step, int local_prop_width, float lambda, float omega,
                           float disp_delta, float delta, int fb_smooth, float s0_ems_prob, float no_change_prob,
                           float range_factor, int update_rigidness_only) {
    return optimize_depth_gpu(h_flows, h_rigidnesses, h_o_rigidnesses, h_depth_priors, h_depth_prior_pconfs,
                              h_depth_prior_confs, h_o_depth_prior_confs, h_depth, h_o_depth, h_K, h_Rs, h_ts, h_dp_Rs,
                              h_dp_ts, abs_resize_factor, N, N_dp, w, h, basefocal, n_rand_samples, global_prop_step,
                              local_prop_width, lambda, omega, disp_delta, delta, fb_smooth != 0, s0_ems_prob,
                              no_change_prob, range_factor, update_rigidness_only != 0);
}

int ref_align_frame_init_gpu(float** h_images, float** h_depths, float** h_weights, float* h_K, float vbf, float crw,
                             int N, int w, int h) {
    return align_frame_init_gpu(h_images, h_depths, h_weights, h_K, vbf, crw, N, w, h);
}

int ref_align_frame_eval_gpu(int ref_fid, int tar_fid, const float* h_params_ref, const float* h_params_tar,
                             float* h_o_residual, float* h_o_jacobian, int apply_weights) {
    return align_frame_eval_gpu(ref_fid, tar_fid, h_params_ref, h_params_tar, h_o_residual, h_o_jacobian,
                                apply_weights != 0);
}

}  // extern "C"
