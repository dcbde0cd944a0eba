// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// Glue around the reference's unmodified host sources inside oracle/_ref/libvoldor_host_ref.so:
//   * the kernel entry points of gpu-kernels/gpu_kernels.h:11-60, which voldor.cpp / geometry.cpp call, are defined
//     here as trampolines onto a kernel library chosen at run time (ref_host_bind: dlopen + symbol prefix — the
//     reference's own kernels "ref_", this repository's ABI "vb_", the CPU port "cpu_"), so that ONE build of the
//     reference host code can drive each of them, and a fresh copy of the reference kernels when a test wants a
//     reference process with no call history;
//   * extern "C" doors for ctypes: ref_host_py_voldor_wrapper (= the reference's py_voldor_wrapper,
//     voldor/py_export.cpp:5-78) and ref_host_bootstrap (VOLDOR::init + VOLDOR::bootstrap, voldor.cpp:4-128,151-162,
//     returning the pose and the closed-form depth the window starts from, so that the other side of a comparison
//     can be started from the same state).
#include <dlfcn.h>
#include <iterator>
#include <sstream>
#include <string>
#include "voldor.h"
#include "py_export.h"

namespace {
typedef int (*fn_meanshift)(float*, float, float*, float*, int*, int, int, int, float, int, int, float);
typedef int (*fn_robust)(float*, float*, float*, float, float, float*, int*, int, int, float, int);
typedef int (*fn_collect)(float**, float**, float*, float*, float**, float**, float*, float*, int, int, int, int, float,
                          float, float, float, int);
typedef int (*fn_p3p)(float*, float*, float*, float*, float*, int, int);
typedef int (*fn_depth)(float**, float**, float**, float**, float**, float**, float**, float*, float*, float*, float**,
                        float**, float**, float**, float, int, int, int, int, float, int, int, int, float, float, float,
                        float, int, float, float, float, int);
struct Bound {
    fn_meanshift meanshift = nullptr;
    fn_robust robust = nullptr;
    fn_collect collect = nullptr;
    fn_p3p twist = nullptr, ap3p = nullptr;
    fn_depth depth = nullptr;
    std::string key;
} g;

void need(const void* f) {
    if (!f) {
        fprintf(stderr, "ref host: no kernel library bound (ref_host_bind)\n");
        abort();
    }
}
}  // namespace

int meanshift_gpu(float* h_space, float kernel_var, float* h_io_mean, float* h_o_confidence, int* used_iters,
                  bool use_external_init_mean, int N, int dims, float epsilon, int max_iters, int max_init_trials,
                  float good_init_confidence) {
    need((void*)g.meanshift);
    return g.meanshift(h_space, kernel_var, h_io_mean, h_o_confidence, used_iters, use_external_init_mean, N, dims,
                       epsilon, max_iters, max_init_trials, good_init_confidence);
}
int fit_robust_gaussian(float* h_space, float* h_io_mean, float* h_io_covar, float trunc_sigma, float covar_reg_lambda,
                        float* h_o_density, int* used_iters, int N, int dims, float epsilon, int max_iters) {
    need((void*)g.robust);
    return g.robust(h_space, h_io_mean, h_io_covar, trunc_sigma, covar_reg_lambda, h_o_density, used_iters, N, dims,
                    epsilon, max_iters);
}
int collect_p3p_instances(float* h_flows[], float* h_rigidnesses[], float* h_depth, float* h_K, float* h_Rs[],
                          float* h_ts[], float* h_o_p2_map, float* h_o_p3_map, int N, int w, int h, int active_idx,
                          float rigidness_thresh, float rigidness_sum_thresh, float sample_min_depth,
                          float sample_max_depth, int max_trace_on_flow) {
    need((void*)g.collect);
    return g.collect(h_flows, h_rigidnesses, h_depth, h_K, h_Rs, h_ts, h_o_p2_map, h_o_p3_map, N, w, h, active_idx,
                     rigidness_thresh, rigidness_sum_thresh, sample_min_depth, sample_max_depth, max_trace_on_flow);
}
int solve_batch_p3p_ap3p_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K, int N_pts,
                             int N_poses) {
    need((void*)g.ap3p);
    return g.ap3p(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses);
}
int solve_batch_p3p_lambdatwist_gpu(float* h_p3s, float* h_p2s, float* h_o_rvecs, float* h_o_tvecs, float* h_K,
                                    int N_pts, int N_poses) {
    need((void*)g.twist);
    return g.twist(h_p3s, h_p2s, h_o_rvecs, h_o_tvecs, h_K, N_pts, N_poses);
}
int optimize_depth_gpu(float* h_flows[], float* h_rigidnesses[], float* h_o_rigidnesses[], float* h_depth_priors[],
                       float* h_depth_prior_pconfs[], float* h_depth_prior_confs[], float* h_o_depth_prior_confs[],
                       float* h_depth, float* h_o_depth, float* h_K, float* h_Rs[], float* h_ts[], float* h_dp_Rs[],
                       float* h_dp_ts[], float abs_resize_factor, int N, int N_dp, int w, int h, float basefocal,
                       int n_rand_samples, int global_prop_step, int local_prop_width, float lambda, float omega,
                       float disp_delta, float delta, bool fb_smooth, float s0_ems_prob, float no_change_prob,
                       float range_factor, bool update_rigidness_only) {
    need((void*)g.depth);
    return g.depth(h_flows, h_rigidnesses, h_o_rigidnesses, h_depth_priors, h_depth_prior_pconfs, h_depth_prior_confs,
                   h_o_depth_prior_confs, h_depth, h_o_depth, h_K, h_Rs, h_ts, h_dp_Rs, h_dp_ts, abs_resize_factor, N,
                   N_dp, w, h, basefocal, n_rand_samples, global_prop_step, local_prop_width, lambda, omega, disp_delta,
                   delta, fb_smooth, s0_ems_prob, no_change_prob, range_factor, update_rigidness_only);
}

extern "C" {

int ref_host_bind(const char* lib_path, const char* prefix) {
    const std::string key = std::string(lib_path) + "|" + prefix;
    if (key == g.key) return 0;
    void* h = dlopen(lib_path, RTLD_NOW | RTLD_LOCAL);
    if (!h) {
        fprintf(stderr, "ref host: dlopen(%s): %s\n", lib_path, dlerror());
        return -1;
    }
    const std::string p(prefix);
    Bound b;
    b.meanshift = (fn_meanshift)dlsym(h, (p + "meanshift_gpu").c_str());
    b.robust = (fn_robust)dlsym(h, (p + "fit_robust_gaussian").c_str());
    b.collect = (fn_collect)dlsym(h, (p + "collect_p3p_instances").c_str());
    b.twist = (fn_p3p)dlsym(h, (p + "solve_batch_p3p_lambdatwist_gpu").c_str());
    b.ap3p = (fn_p3p)dlsym(h, (p + "solve_batch_p3p_ap3p_gpu").c_str());
    b.depth = (fn_depth)dlsym(h, (p + "optimize_depth_gpu").c_str());
    if (!(b.meanshift && b.robust && b.collect && b.twist && b.ap3p && b.depth)) return -2;
    b.key = key;
    g = b;
    return 0;
}

int ref_host_py_voldor_wrapper(const float* flows, const float* disparity, const float* disparity_pconf,
                               const float* depth_priors, const float* depth_prior_poses,
                               const float* depth_prior_pconfs, float fx, float fy, float cx, float cy, float basefocal,
                               int N, int N_dp, int w, int h, const char* config, int* n_registered, float* poses,
                               float* poses_covar, float* depth, float* depth_conf) {
    int n = 0;
    const int rc = py_voldor_wrapper(flows, disparity, disparity_pconf, depth_priors, depth_prior_poses,
                                     depth_prior_pconfs, fx, fy, cx, cy, basefocal, N, N_dp, w, h, config, n, poses,
                                     poses_covar, depth, depth_conf);
    *n_registered = n;
    return rc;
}

// the state a monocular window starts from: camera 0 after estimate_camera_pose_epipolar (pose injected through
// cvmin_inject_epipolar, then rotated by the reference's own "t = R t") and the reference's closed-form depth
int ref_host_bootstrap(const float* flows_pt, float fx, float fy, float cx, float cy, int N, int w, int h,
                       const char* config_pt, float* R9, float* t3, float* depth_out) {
    Config cfg;
    std::istringstream iss(config_pt);
    std::vector<std::string> strs(std::istream_iterator<std::string>{iss}, std::istream_iterator<std::string>());
    cfg.fx = fx, cfg.cx = cx, cfg.fy = fy, cfg.cy = cy;
    cfg.read_config(strs);
    std::vector<Mat> flows;
    for (int i = 0; i < N; i++) flows.push_back(Mat(Size(w, h), CV_32FC2, (void*)(flows_pt + (size_t)i * w * h * 2)));
    VOLDOR v(cfg);
    v.init(flows);
    v.bootstrap();
    memcpy(R9, v.cams[0].R.data, 9 * sizeof(float));
    memcpy(t3, v.cams[0].t.data, 3 * sizeof(float));
    memcpy(depth_out, v.depth.data, (size_t)w * h * sizeof(float));
    return 0;
}

}  // extern "C"
