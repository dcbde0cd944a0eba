// voldor_b200 — drop-in boundary of the EM hot path (library level).
//
// This header declares, with C++ linkage and the exact signatures (names, parameter order, default
// arguments) of the reference's gpu-kernels/gpu_kernels.h:11-74, the eight entry points that the
// reference's host orchestration (voldor/voldor.cpp, voldor/geometry.cpp, frame-alignment/*.h) links
// against, so that those translation units compile and link unchanged against libvoldor_b200.so.
// The same functions are also exported with C linkage under the prefix vb_ (see voldor_b200.h) for
// dlopen/ctypes/cgo-style FFI users; bool parameters become int there.
//
// Contract kept from the reference:
//   * every pointer is a HOST pointer owned by the caller; arrays of pointers are tables of row-major images;
//   * a NULL input pointer means "reuse the device copy cached by the previous call of the same function
//     family", a NULL output pointer means "skip that download";
//   * return value: 0 (cudaSuccess) or the cudaError_t after printing "GPUassert : ..."; fit_robust_gaussian
//     returns 1 when the fit is unreliable and leaves its outputs untouched;
//   * entry points are serialised internally (one mutex per family); device state is per process.
#pragma once

#if defined(WIN32) || defined(_WIN32)
#define DLL_EXPORT __declspec(dllexport)
#else
#define DLL_EXPORT __attribute__((visibility("default")))
#endif

// replaces reference gpu-kernels/meanshift.cu:34-150 (declared gpu_kernels.h:11-15)
DLL_EXPORT int meanshift_gpu(float* h_space, float kernel_var,
	float* h_io_mean, float* h_o_confidence, int* used_iters,
	bool use_external_init_mean, int N, int dims,
	float epsilon = 1e-5f, int max_iters = 100,
	int max_init_trials = 20, float good_init_confidence = 0.5f);

// replaces reference gpu-kernels/fit_robust_gaussian.cu:101-286 (declared gpu_kernels.h:17-22)
DLL_EXPORT int fit_robust_gaussian(
	float* h_space, float* h_io_mean, float* h_io_covar,
	float trunc_sigma, float covar_reg_lambda,
	float* h_o_density, int* used_iters,
	int N, int dims,
	float epsilon, int max_iters);

// replaces reference gpu-kernels/collect_p3p_instances.cu:147-250 (declared gpu_kernels.h:24-35)
DLL_EXPORT int collect_p3p_instances(
	float* h_flows[], float* h_rigidnesses[],
	float* h_depth,
	float* h_K, float* h_Rs[], float* h_ts[],
	float* h_o_p2_map, float* h_o_p3_map,
	int N, int w, int h,
	int active_idx,
	float rigidness_thresh,
	float rigidness_sum_thresh,
	float sample_min_depth,
	float sample_max_depth,
	int max_trace_on_flow);

// replaces reference gpu-kernels/solve_batch_ap3p.cu:387-437 (declared gpu_kernels.h:37-39)
DLL_EXPORT int solve_batch_p3p_ap3p_gpu(float* h_p3s, float* h_p2s,
	float* h_o_rvecs, float* h_o_tvecs,
	float* h_K, int N_pts, int N_poses);
// replaces reference gpu-kernels/solve_batch_lambdatwist.cu:51-102 (declared gpu_kernels.h:40-42)
DLL_EXPORT int solve_batch_p3p_lambdatwist_gpu(float* h_p3s, float* h_p2s,
	float* h_o_rvecs, float* h_o_tvecs,
	float* h_K, int N_pts, int N_poses);

// replaces reference gpu-kernels/optimize_depth.cu:293-520 (declared gpu_kernels.h:44-58)
DLL_EXPORT int optimize_depth_gpu(
	float* h_flows[],
	float* h_rigidnesses[], float* h_o_rigidnesses[],
	float* h_depth_priors[], float* h_depth_prior_pconfs[],
	float* h_depth_prior_confs[], float* h_o_depth_prior_confs[],
	float* h_depth, float* h_o_depth,
	float* h_K, float* h_Rs[], float* h_ts[],
	float* h_dp_Rs[], float* h_dp_ts[],
	float abs_resize_factor,
	int N, int N_dp, int w, int h, float basefocal,
	int n_rand_samples, int global_prop_step, int local_prop_width,
	float lambda, float omega, float disp_delta, float delta,
	bool fb_smooth, float s0_ems_prob, float no_change_prob,
	float range_factor,
	bool update_rigidness_only);

// replaces reference gpu-kernels/align_frame.cu:512-554 (declared gpu_kernels.h:60-66)
DLL_EXPORT int align_frame_init_gpu(
	float* h_images[],
	float* h_depths[],
	float* h_weights[],
	float* h_K,
	float vbf, float crw,
	int N, int w, int h);

// replaces reference gpu-kernels/align_frame.cu:414-510 (declared gpu_kernels.h:68-74)
DLL_EXPORT int align_frame_eval_gpu(
	int ref_fid,
	int tar_fid,
	const float* h_params_ref,
	const float* h_params_tar,
	float* h_o_residual, float* h_o_jacobian,
	const bool apply_weights = true);
