// voldor_b200 — drop-in boundary of the EM hot path (library level).
//
// The eight entry points below carry, with C++ linkage, the signatures of the reference's
// gpu-kernels/gpu_kernels.h:11-74 (same function names, parameter types and order, same default arguments), so
// the reference's host translation units (voldor/voldor.cpp, voldor/geometry.cpp, frame-alignment/*.h) compile
// and link unchanged against libvoldor_b200.so; tests/test_cpu_abi_symbols.py checks the mangled names.  Parameter
// names and documentation are this project's.  The same functions are exported with C linkage under the prefix
// vb_ (voldor_b200.h) for dlopen / ctypes / cgo-style users; bool parameters become int there.
//
// Contract kept from the reference:
//   * every pointer may be a HOST pointer owned by the caller (device pointers are accepted too); "table" means a
//     caller-allocated array of pointers to row-major images;
//   * a NULL input means "reuse the device copy cached by the previous call of the same function family", a NULL
//     output means "skip that download";
//   * return value: 0 (cudaSuccess) or the cudaError_t after printing "GPUassert : ..."; fit_robust_gaussian
//     returns 1 when the fit is unreliable and leaves its outputs untouched;
//   * entry points are serialised internally; device state is per process.
#pragma once

#if defined(WIN32) || defined(_WIN32)
#define DLL_EXPORT __declspec(dllexport)
#else
#define DLL_EXPORT __attribute__((visibility("default")))
#endif

// Gaussian-kernel mean-shift mode of a point cloud.
// Replaces reference gpu-kernels/meanshift.cu:34-150 (declared gpu_kernels.h:11-15).
DLL_EXPORT int meanshift_gpu(
    float* points,                       // [count][dimension], row-major
    float kernel_variance,               // isotropic Gaussian kernel variance
    float* mean_inout,                   // [dimension] start point if seeded_start, receives the mode
    float* confidence_out,               // sum of kernel weights at the mode / count
    int* iterations_out,                 //
    bool seeded_start,                   // false: best of <= start_trials random samples (libc rand())
    int count,                           //
    int dimension,                       // <= 16
    float tolerance = 1e-5f,             // stop when the mean moves less than this
    int iteration_limit = 100,           //
    int start_trials = 20,               //
    float good_start_fraction = 0.5f);   // a trial whose weight sum exceeds this fraction of count ends the trials

// Truncated-Gaussian EM (hard 0/1 weights inside trunc_sigma), 6-D at most.
// Replaces reference gpu-kernels/fit_robust_gaussian.cu:101-286 (declared gpu_kernels.h:17-22).
DLL_EXPORT int fit_robust_gaussian(
    float* points,                       // [count][dimension]
    float* mean_inout,                   // [dimension]
    float* covariance_inout,             // [dimension][dimension], full symmetric
    float truncation_sigma,              // Mahalanobis gate
    float shrinkage_lambda,              // Ledoit-Wolf shrink towards tr/n * I from the 2nd iteration on
    float* density_out,                  // fraction of points inside the gate
    int* iterations_out,                 //
    int count,                           //
    int dimension,                       //
    float tolerance,                     // on the change of the density
    int iteration_limit);

// Per-pixel P3P instances (3-D point before motion `camera`, 2-D observation after it), NaN where invalid.
// Replaces reference gpu-kernels/collect_p3p_instances.cu:147-250 (declared gpu_kernels.h:24-35).
DLL_EXPORT int collect_p3p_instances(
    float* flow_table[],                 // [frames] x [rows][cols][2]
    float* rigidness_table[],            // [frames] x [rows][cols]
    float* depth,                        // [rows][cols]
    float* intrinsics,                   // 3x3 row-major
    float* rotation_table[],             // [frames] x 3x3
    float* translation_table[],          // [frames] x 3
    float* points2d_out,                 // [rows][cols][2]
    float* points3d_out,                 // [rows][cols][3]
    int frames,                          //
    int cols,                            //
    int rows,                            //
    int camera,                          // index of the motion to be estimated
    float rigidness_gate,                // on the product of traced rigidness values
    float rigidness_sum_gate,            //
    float min_depth,                     //
    float max_depth,                     //
    int max_flow_trace);                 // how many frames the observation follows the measured flow

// Batched P4P pose hypotheses from random quadruples (XORWOW seed 233): AP3P and lambdatwist solvers.
// Replace reference gpu-kernels/solve_batch_ap3p.cu:387-437 and solve_batch_lambdatwist.cu:51-102
// (declared gpu_kernels.h:37-42).
DLL_EXPORT int solve_batch_p3p_ap3p_gpu(
    float* points3d,                     // [point_count][3]
    float* points2d,                     // [point_count][2]
    float* rvecs_out,                    // [hypotheses][3], NaN for a failed hypothesis
    float* tvecs_out,                    // [hypotheses][3]
    float* intrinsics,                   // 3x3 row-major
    int point_count,                     //
    int hypotheses);
DLL_EXPORT int solve_batch_p3p_lambdatwist_gpu(
    float* points3d,
    float* points2d,
    float* rvecs_out,
    float* tvecs_out,
    float* intrinsics,
    int point_count,
    int hypotheses);

// One depth M-step (random search + propagation) and rigidness E-step of a window.
// Replaces reference gpu-kernels/optimize_depth.cu:293-520 (declared gpu_kernels.h:44-58).
DLL_EXPORT int optimize_depth_gpu(
    float* flow_table[],                 // [frames] x [rows][cols][2]
    float* rigidness_table[],            // in:  [frames] x [rows][cols]
    float* rigidness_out_table[],        // out: same shape (may alias the input table)
    float* prior_table[],                // [priors] x [rows][cols] depth priors
    float* prior_pixel_conf_table[],     // [priors] x per-pixel confidence supplied with the prior
    float* prior_conf_table[],           // in:  [priors] x posterior confidence of the prior
    float* prior_conf_out_table[],       // out: same
    float* depth,                        // in:  [rows][cols]
    float* depth_out,                    // out: [rows][cols] (may alias)
    float* intrinsics,                   // 3x3
    float* rotation_table[],             // [frames] x 3x3
    float* translation_table[],          // [frames] x 3
    float* prior_rotation_table[],       // [priors] x 3x3
    float* prior_translation_table[],    // [priors] x 3
    float absolute_resize_factor,        //
    int frames,                          //
    int priors,                          //
    int cols,                            //
    int rows,                            //
    float basefocal,                     // stereo baseline * focal length (0 for monocular)
    int random_samples,                  //
    int global_propagation_step,         //
    int local_propagation_width,         //
    float lambda,                        // flow outlier scale
    float omega,                         // prior outlier scale
    float disparity_delta,               // weight of a disparity prior (prior 0) when > 0
    float delta,                         // weight of a depth prior
    bool smooth_rigidness,               // forward-backward HMM smoothing before the M-step
    float state0_emission,               //
    float no_change_probability,         //
    float range_factor,                  // of the inverse-depth sampling range
    bool rigidness_only);                // skip the M-step

// Frame alignment: upload N frames (init) and evaluate residuals + Jacobian between two of them (eval).
// Replace reference gpu-kernels/align_frame.cu:512-554 and :414-510 (declared gpu_kernels.h:60-74).
DLL_EXPORT int align_frame_init_gpu(
    float* image_table[],                // [frames] x [rows][cols] grey images
    float* depth_table[],                // [frames] x [rows][cols]
    float* weight_table[],               // [frames] x [rows][cols]
    float* intrinsics,                   // 3x3
    float virtual_basefocal,             //
    float colour_residual_weight,        //
    int frames,                          //
    int cols,                            //
    int rows);
DLL_EXPORT int align_frame_eval_gpu(
    int reference_frame,                 //
    int target_frame,                    //
    const float* reference_params,       // [9] rvec, tvec, log depth scale, colour scale, colour offset
    const float* target_params,          // [9]
    float* residual_out,                 // [rows][cols]
    float* jacobian_out,                 // [rows][cols][9]
    const bool weighted_loss = true);
