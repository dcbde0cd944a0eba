// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// OpenCV-free restatement of the reference's HOST orchestration of one VO window, written against the
// library-level ABI (gpu_kernels.h) exactly the way the reference drives it: host buffers in and out of
// every call, NULL = "use the cached device copy", host-side compaction / filtering / scaling between calls.
//   voldor/py_export.cpp:5-79      -> oracle_py_voldor_wrapper
//   voldor/voldor.cpp:4-128        -> HostWindow::init
//   voldor/voldor.cpp:130-149      -> HostWindow::solve
//   voldor/voldor.cpp:164-201      -> HostWindow::optimize_cameras
//   voldor/voldor.cpp:203-307      -> HostWindow::optimize_depth  (both call patterns :250-290)
//   voldor/voldor.cpp:309-317      -> HostWindow::normalize_world_scale
//   voldor/geometry.cpp:5-265      -> HostWindow::optimize_camera_pose
// The kernels behind the ABI are chosen at run time (dlopen + symbol prefix):
//   oracle/_ref/libgpu_kernels_ref.so (prefix ref_)  = the reference's own CUDA kernels for sm_100a
//   voldor_b200/libvoldor_b200.so     (prefix vb_)   = our kernels through the same host-pointer ABI
//   oracle/libvoldor_oracle.so        (prefix cpu_)  = the CPU port (cpu_kernels.cpp), used as cpu_baseline
// so the window-level parity tests compare our device-resident pipeline against "reference kernels +
// reference orchestration".  Third-party arithmetic of the reference (cv::Rodrigues, cv::norm, Mat scaling)
// is restated in ../voldor_b200/csrc/host_math.h; the monocular bootstrap (cv::findEssentialMat) is injected
// by the caller.  Flag grammar: ../voldor_b200/csrc/config.h.
#include "../voldor_b200/csrc/config.h"
#include "../voldor_b200/csrc/host_math.h"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <string>
#include <vector>

namespace {

typedef int (*fn_meanshift)(float*, float, float*, float*, int*, int, int, int, float, int, int, float);
typedef int (*fn_robust)(float*, float*, float*, float, float, float*, int*, int, int, float, int);
typedef int (*fn_collect)(float**, float**, float*, float*, float**, float**, float*, float*, int, int, int, int, float,
                          float, float, float, int);
typedef int (*fn_p3p)(float*, float*, float*, float*, float*, int, int);
typedef int (*fn_depth)(float**, float**, float**, float**, float**, float**, float**, float*, float*, float*, float**,
                        float**, float**, float**, float, int, int, int, int, float, int, int, int, float, float, float,
                        float, int, float, float, float, int);

struct Abi {
    void* handle = nullptr;
    fn_meanshift meanshift = nullptr;
    fn_robust robust = nullptr;
    fn_collect collect = nullptr;
    fn_p3p p3p_lambdatwist = nullptr, p3p_ap3p = nullptr;
    fn_depth depth = nullptr;
    bool load(const char* path, const char* prefix) {
        handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
        if (!handle) {
            printf("oracle host: dlopen(%s) failed: %s\n", path, dlerror());
            return false;
        }
        std::string p(prefix);
        meanshift = (fn_meanshift)dlsym(handle, (p + "meanshift_gpu").c_str());
        robust = (fn_robust)dlsym(handle, (p + "fit_robust_gaussian").c_str());
        collect = (fn_collect)dlsym(handle, (p + "collect_p3p_instances").c_str());
        p3p_lambdatwist = (fn_p3p)dlsym(handle, (p + "solve_batch_p3p_lambdatwist_gpu").c_str());
        p3p_ap3p = (fn_p3p)dlsym(handle, (p + "solve_batch_p3p_ap3p_gpu").c_str());
        depth = (fn_depth)dlsym(handle, (p + "optimize_depth_gpu").c_str());
        return meanshift && robust && collect && p3p_lambdatwist && p3p_ap3p && depth;
    }
};

struct Camera {
    float K[9];
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float t[3] = {0, 0, 0};
    float pose_covar[36] = {0};
    float pose_density = 0;
    int pose_sample_count = 0;
    float pose_rigidness_density = 0;
    int last_used_ms_iters = 0, last_used_gu_iters = 0;
};

struct HostWindow {
    Abi* abi;
    vb::Config cfg;
    int w = 0, h = 0, n_flows = 0, n_flows_init = 0, n_depth_priors = 0, iters_cur = 0, iters_remain = 0;
    bool has_disparity = false;
    std::vector<std::vector<float>> flows, rigidnesses, depth_priors, depth_prior_pconfs, depth_prior_confs;
    std::vector<float> depth;
    std::vector<Camera> cams, prior_cams;
    double t_cameras = 0, t_depth = 0;

    bool init(const float* flows_pt, const float* disparity_pt, const float* disparity_pconf_pt,
              const float* depth_priors_pt, const float* depth_prior_poses_pt, const float* depth_prior_pconfs_pt,
              int N, int N_dp) {
        const size_t npx = (size_t)w * h;
        iters_cur = 0, iters_remain = cfg.max_iters;
        if (cfg.resize_factor != 1) return false;
        for (int i = 0; i < N; i++) flows.emplace_back(flows_pt + i * npx * 2, flows_pt + (i + 1) * npx * 2);
        if (disparity_pt) {
            std::vector<float> dp(npx);
            for (size_t k = 0; k < npx; k++) dp[k] = disparity_pt[k] == 0.f ? 0.f : cfg.basefocal / disparity_pt[k];
            depth_priors.push_back(dp);
            if (disparity_pconf_pt)
                depth_prior_pconfs.emplace_back(disparity_pconf_pt, disparity_pconf_pt + npx);
            else
                depth_prior_pconfs.emplace_back(npx, 1.f);
            prior_cams.emplace_back();
        }
        for (int i = 0; i < N_dp; i++) {
            depth_priors.emplace_back(depth_priors_pt + i * npx, depth_priors_pt + (i + 1) * npx);
            if (depth_prior_pconfs_pt)
                depth_prior_pconfs.emplace_back(depth_prior_pconfs_pt + i * npx, depth_prior_pconfs_pt + (i + 1) * npx);
            else
                depth_prior_pconfs.emplace_back(npx, 1.f);
            Camera cam;
            vb::hm::rvec_to_matrix(depth_prior_poses_pt + i * 6, cam.R);
            memcpy(cam.t, depth_prior_poses_pt + i * 6 + 3, 3 * sizeof(float));
            prior_cams.push_back(cam);
        }
        n_flows = n_flows_init = (int)flows.size();
        n_depth_priors = (int)depth_priors.size();
        has_disparity = disparity_pt != nullptr;
        for (int i = 0; i < n_depth_priors; i++) depth_prior_confs.emplace_back(npx, 1.f);
        for (int i = 0; i < n_flows; i++) rigidnesses.emplace_back(npx, 1.f);
        for (int i = 0; i < n_flows; i++) {
            Camera cam;
            const float K[9] = {cfg.fx, 0, cfg.cx, 0, cfg.fy, cfg.cy, 0, 0, 1};
            memcpy(cam.K, K, sizeof(K));
            cams.push_back(cam);
        }
        if (n_depth_priors > 0) {
            depth = depth_priors[0];
            if (!disparity_pt) optimize_depth(1);
        } else {
            depth.assign(npx, 1.f);
        }
        return true;
    }

    // flag: 0 default, 1 only depth priors, 2 rigidness only (voldor.h:7-11)
    void optimize_depth(int flag) {
        if (n_flows == 0 && n_depth_priors == 0) return;
        std::vector<float*> h_flows, h_rig, h_Rs, h_ts, h_dp, h_pc, h_cf, h_dR, h_dt;
        const bool with_flows = n_flows > 0 && flag != 1;
        if (with_flows)
            for (int i = 0; i < n_flows; i++) {
                h_flows.push_back(flows[i].data()), h_rig.push_back(rigidnesses[i].data());
                h_Rs.push_back(cams[i].R), h_ts.push_back(cams[i].t);
            }
        for (int i = 0; i < n_depth_priors; i++) {
            h_dp.push_back(depth_priors[i].data()), h_pc.push_back(depth_prior_pconfs[i].data());
            h_cf.push_back(depth_prior_confs[i].data());
            h_dR.push_back(prior_cams[i].R), h_dt.push_back(prior_cams[i].t);
        }
        float** pf = with_flows ? h_flows.data() : nullptr;
        float** pr = with_flows ? h_rig.data() : nullptr;
        float** pR = with_flows ? h_Rs.data() : nullptr;
        float** pt = with_flows ? h_ts.data() : nullptr;
        float** pdp = n_depth_priors ? h_dp.data() : nullptr;
        float** ppc = n_depth_priors ? h_pc.data() : nullptr;
        float** pcf = n_depth_priors ? h_cf.data() : nullptr;
        float** pdR = n_depth_priors ? h_dR.data() : nullptr;
        float** pdt = n_depth_priors ? h_dt.data() : nullptr;
        const float disp_delta = has_disparity ? cfg.disp_delta : -1;
        if (!cfg.exclusive_gpu_context || iters_cur == 0 || iters_cur == 1) {
            abi->depth(pf, pr, pr, pdp, ppc, pcf, pcf, depth.data(), depth.data(), cams[0].K, pR, pt, pdR, pdt,
                       cfg.abs_resize_factor, flag == 1 ? 0 : n_flows, n_depth_priors, w, h, cfg.basefocal,
                       cfg.depth_rand_samples, cfg.depth_global_prop_step, cfg.depth_local_prop_width, cfg.lambda,
                       cfg.omega, disp_delta, cfg.delta, cfg.fb_smooth, cfg.fb_emm, cfg.fb_no_change_prob,
                       cfg.depth_range_factor, flag == 2);
        } else {
            abi->depth(nullptr, nullptr, pr, nullptr, nullptr, nullptr, pcf, nullptr, depth.data(), nullptr, pR, pt,
                       nullptr, nullptr, cfg.abs_resize_factor, n_flows, n_depth_priors, w, h, cfg.basefocal,
                       cfg.depth_rand_samples, cfg.depth_global_prop_step, cfg.depth_local_prop_width, cfg.lambda,
                       cfg.omega, disp_delta, cfg.delta, cfg.fb_smooth, cfg.fb_emm, cfg.fb_no_change_prob,
                       cfg.depth_range_factor, flag == 2);
        }
    }

    int optimize_camera_pose(int active_idx, bool successive_pose, bool rg_refine, bool update_batch_instance,
                             bool update_iter_instance) {
        const size_t npx = (size_t)w * h;
        std::vector<float*> h_flows, h_rig, h_Rs, h_ts;
        for (int i = 0; i < n_flows; i++) {
            h_flows.push_back(flows[i].data()), h_rig.push_back(rigidnesses[i].data());
            h_Rs.push_back(cams[i].R), h_ts.push_back(cams[i].t);
        }
        std::vector<float> p2_map(npx * 2), p3_map(npx * 3);
        if (update_batch_instance)
            abi->collect(h_flows.data(), h_rig.data(), depth.data(), cams[0].K, h_Rs.data(), h_ts.data(), p2_map.data(),
                         p3_map.data(), n_flows, w, h, active_idx, cfg.rigidness_threshold,
                         cfg.rigidness_sum_threshold, cfg.pose_sample_min_depth, cfg.pose_sample_max_depth,
                         cfg.max_trace_on_flow);
        else if (update_iter_instance)
            abi->collect(nullptr, h_rig.data(), depth.data(), nullptr, h_Rs.data(), h_ts.data(), p2_map.data(),
                         p3_map.data(), n_flows, w, h, active_idx, cfg.rigidness_threshold,
                         cfg.rigidness_sum_threshold, cfg.pose_sample_min_depth, cfg.pose_sample_max_depth,
                         cfg.max_trace_on_flow);
        else
            abi->collect(nullptr, nullptr, nullptr, nullptr, h_Rs.data(), h_ts.data(), p2_map.data(), p3_map.data(),
                         n_flows, w, h, active_idx, cfg.rigidness_threshold, cfg.rigidness_sum_threshold,
                         cfg.pose_sample_min_depth, cfg.pose_sample_max_depth, cfg.max_trace_on_flow);

        // raster-order compaction of finite instances (geometry.cpp:66-80)
        std::vector<float> pts2(npx * 2 + 2), pts3(npx * 3 + 3);
        int n_points = 0;
        for (size_t i = 0; i < npx; i++) {
            const float* a = &p2_map[i * 2];
            const float* b = &p3_map[i * 3];
            if (std::isfinite(a[0] + a[1] + b[0] + b[1] + b[2])) {
                pts2[n_points * 2] = a[0], pts2[n_points * 2 + 1] = a[1];
                pts3[n_points * 3] = b[0], pts3[n_points * 3 + 1] = b[1], pts3[n_points * 3 + 2] = b[2];
                n_points++;
            }
        }
        if (n_points < 4) return 0;

        const int P = cfg.n_poses_to_sample;
        std::vector<float> pool((size_t)P * 6);
        int used = 0;
        {
            std::vector<float> ret_R((size_t)P * 3), ret_t((size_t)P * 3);
            if (cfg.lambdatwist)
                abi->p3p_lambdatwist(pts3.data(), pts2.data(), ret_R.data(), ret_t.data(), cams[active_idx].K, n_points, P);
            else
                abi->p3p_ap3p(pts3.data(), pts2.data(), ret_R.data(), ret_t.data(), cams[active_idx].K, n_points, P);
            for (int i = 0; i < P; i++) {
                if (std::isfinite(ret_R[i * 3] + ret_R[i * 3 + 1] + ret_R[i * 3 + 2] + ret_t[i * 3] + ret_t[i * 3 + 1] +
                                  ret_t[i * 3 + 2])) {
                    for (int d = 0; d < 3; d++) pool[used * 6 + d] = ret_R[i * 3 + d], pool[used * 6 + 3 + d] = ret_t[i * 3 + d];
                    used++;
                }
            }
        }
        if (used == 0) return 0;
        Camera& cam = cams[active_idx];
        cam.pose_sample_count = used;

        float pose_opm[6];
        vb::hm::matrix_to_rvec(cam.R, pose_opm);
        memcpy(pose_opm + 3, cam.t, 3 * sizeof(float));
        for (int i = 0; i < used; i++)
            for (int d = 0; d < 3; d++) pool[i * 6 + d] *= cfg.meanshift_rvec_scale;
        for (int d = 0; d < 3; d++) pose_opm[d] *= cfg.meanshift_rvec_scale;
        abi->meanshift(pool.data(), cfg.meanshift_kernel_var, pose_opm, &cam.pose_density, &cam.last_used_ms_iters,
                       successive_pose, used, 6, cfg.meanshift_epsilon, cfg.meanshift_max_iters,
                       cfg.meanshift_max_init_trials, cfg.meanshift_good_init_confidence);

        if (rg_refine) {
            const float sc2 = cfg.rg_pose_scaling * cfg.rg_pose_scaling;
            for (int k = 0; k < 36; k++) cam.pose_covar[k] = 0;
            for (int d = 0; d < 6; d++) cam.pose_covar[d * 6 + d] = cfg.meanshift_kernel_var;
            for (int k = 0; k < 36; k++) cam.pose_covar[k] *= sc2;
            for (int d = 0; d < 6; d++) pose_opm[d] *= cfg.rg_pose_scaling;
            for (size_t k = 0; k < (size_t)used * 6; k++) pool[k] *= cfg.rg_pose_scaling;
            const int ret = abi->robust(pool.data(), pose_opm, cam.pose_covar, cfg.rg_trunc_sigma, cfg.rg_covar_reg_lambda,
                                        &cam.pose_density, &cam.last_used_gu_iters, used, 6, cfg.rg_epsilon,
                                        cfg.rg_max_iters);
            if (ret == 0) {
                const float inv = (float)(1. / (double)sc2);
                for (int k = 0; k < 36; k++) cam.pose_covar[k] *= inv;
                for (int i1 = 0; i1 < 6; i1++)
                    for (int i2 = 0; i2 < 6; i2++) {
                        if (i1 < 3 || i2 < 3) cam.pose_covar[i1 * 6 + i2] /= cfg.meanshift_rvec_scale;
                        if (i1 < 3 && i2 < 3) cam.pose_covar[i1 * 6 + i2] /= cfg.meanshift_rvec_scale;
                    }
            } else {
                for (int k = 0; k < 36; k++) cam.pose_covar[k] = 0;
            }
            const float invs = (float)(1. / (double)cfg.rg_pose_scaling);
            for (int d = 0; d < 6; d++) pose_opm[d] *= invs;
        }
        const float invr = (float)(1. / (double)cfg.meanshift_rvec_scale);
        for (int d = 0; d < 3; d++) pose_opm[d] *= invr;
        for (int d = 0; d < 6; d++)
            if (!std::isfinite(pose_opm[d])) return 0;
        vb::hm::rvec_to_matrix(pose_opm, cam.R);
        memcpy(cam.t, pose_opm + 3, 3 * sizeof(float));
        return 1;
    }

    void optimize_cameras() {
        const bool allow_trunc = iters_cur > cfg.no_trunc_iters;
        for (int i = 0; i < n_flows; i++) {
            double sum = 0;
            for (float v : rigidnesses[i]) sum += v;
            cams[i].pose_rigidness_density = (float)sum / (float)(w * h);
            int ok = 0;
            if (!allow_trunc || cams[i].pose_rigidness_density > cfg.trunc_rigidness_density)
                ok = optimize_camera_pose(i, cams[i].pose_sample_count != 0,
                                          cfg.rg_refine && (!cfg.rg_refine_last_only || iters_remain == 0),
                                          !cfg.exclusive_gpu_context || (iters_cur == 1 && i == 0), i == 0);
            if (!ok || (allow_trunc && cams[i].pose_density < cfg.trunc_sample_density)) {
                iters_remain = std::max(iters_remain, cfg.min_iters_after_trunc);
                n_flows = i;
                break;
            }
        }
    }

    void normalize_world_scale() {
        float world_scale = 0;
        for (int i = 0; i < n_flows; i++) world_scale = (float)((double)world_scale + vb::hm::norm3(cams[i].t));
        const float sc = n_flows / world_scale;
        for (int i = 0; i < n_flows; i++)
            for (int d = 0; d < 3; d++) cams[i].t[d] *= sc;
        for (float& v : depth) v *= sc;
    }

    int solve() {
        using clk = std::chrono::high_resolution_clock;
        while (iters_remain > 0 && n_flows > 0) {
            iters_cur++, iters_remain--;
            auto t0 = clk::now();
            optimize_cameras();
            auto t1 = clk::now();
            optimize_depth(cfg.optimize_depth ? 0 : 2);
            if (cfg.norm_world_scale && n_depth_priors == 0) normalize_world_scale();
            auto t2 = clk::now();
            t_cameras += std::chrono::duration<double, std::milli>(t1 - t0).count();
            t_depth += std::chrono::duration<double, std::milli>(t2 - t1).count();
        }
        return iters_cur;
    }
};

Abi g_abi;
std::string g_abi_key;

}  // namespace

extern "C" {

// choose the kernel library that sits behind the ABI
int oracle_host_bind(const char* lib_path, const char* prefix) {
    const std::string key = std::string(lib_path) + "|" + prefix;
    if (key == g_abi_key) return 0;
    Abi a;
    if (!a.load(lib_path, prefix)) return -1;
    g_abi = a;
    g_abi_key = key;
    return 0;
}

// py_voldor_wrapper restated over the bound ABI.  boot_* (optional): injected monocular bootstrap
// (pose of camera 0 + depth map) replacing cv::findEssentialMat/recoverPose + closed-form depth.
int oracle_py_voldor_wrapper(const float* flows_pt, const float* disparity_pt, const float* disparity_pconf_pt,
                             const float* depth_priors_pt, const float* depth_prior_poses_pt,
                             const float* depth_prior_pconfs_pt, float fx, float fy, float cx, float cy,
                             float basefocal, int N, int N_dp, int w, int h, const char* config_pt,
                             const float* boot_R9, const float* boot_t3, const float* boot_depth, int* n_registered,
                             float* poses_pt, float* poses_covar_pt, float* depth_pt, float* depth_conf_pt,
                             int* iters_run, float* stats) {
    auto t0 = std::chrono::high_resolution_clock::now();
    HostWindow W;
    W.abi = &g_abi;
    W.cfg.fx = fx, W.cfg.cx = cx, W.cfg.fy = fy, W.cfg.cy = cy, W.cfg.basefocal = basefocal;
    W.cfg.read(config_pt);
    W.w = w, W.h = h;
    *n_registered = 0;
    if (!W.init(flows_pt, disparity_pt, disparity_pconf_pt, depth_priors_pt, depth_prior_poses_pt, depth_prior_pconfs_pt,
                N, N_dp))
        return 0;
    if (W.n_depth_priors == 0) {
        if (!boot_depth) {
            printf("oracle host: monocular window needs an injected bootstrap\n");
            return 0;
        }
        memcpy(W.cams[0].R, boot_R9, 9 * sizeof(float));
        memcpy(W.cams[0].t, boot_t3, 3 * sizeof(float));
        W.depth.assign(boot_depth, boot_depth + (size_t)w * h);
    }
    W.solve();
    *n_registered = W.n_flows;
    const size_t npx = (size_t)w * h;
    for (int i = 0; i < W.n_flows; i++) {
        if (poses_pt) {
            vb::hm::matrix_to_rvec(W.cams[i].R, poses_pt + i * 6);
            memcpy(poses_pt + i * 6 + 3, W.cams[i].t, 3 * sizeof(float));
        }
        if (poses_covar_pt) memcpy(poses_covar_pt + i * 36, W.cams[i].pose_covar, 36 * sizeof(float));
    }
    if (depth_pt) memcpy(depth_pt, W.depth.data(), npx * sizeof(float));
    if (depth_conf_pt) {
        std::vector<float> conf(npx, 0.f);
        for (int i = 0; i < W.n_flows; i++)
            for (size_t k = 0; k < npx; k++) conf[k] += W.rigidnesses[i][k];
        for (int i = 0; i < W.n_depth_priors; i++)
            for (size_t k = 0; k < npx; k++) conf[k] += W.depth_prior_confs[i][k];
        const float inv_n = (float)(1. / (double)(float)(W.n_flows + W.n_depth_priors));
        for (size_t k = 0; k < npx; k++) depth_conf_pt[k] = conf[k] * inv_n;
    }
    if (iters_run) *iters_run = W.iters_cur;
    if (stats) {
        stats[0] = (float)std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        stats[1] = (float)W.t_cameras, stats[2] = (float)W.t_depth, stats[3] = 0;
    }
    return 0;
}

}  // extern "C"

// ---- small probes used by the GPU-less tests (host logic shared with the product: config grammar, rotations)
extern "C" {

void oracle_rvec_to_matrix(const float* rvec, float* R9) { vb::hm::rvec_to_matrix(rvec, R9); }
void oracle_matrix_to_rvec(const float* R9, float* rvec) { vb::hm::matrix_to_rvec(R9, rvec); }

// parse a flag string on top of the defaults; returns a few representative fields
void oracle_config_probe(const char* cfg_str, float* out /*[12]*/) {
    vb::Config c;
    c.read(cfg_str);
    out[0] = (float)c.max_iters, out[1] = c.no_trunc_iters, out[2] = (float)c.n_poses_to_sample;
    out[3] = c.silent ? 1.f : 0.f, out[4] = c.lambda, out[5] = (float)c.lambdatwist, out[6] = c.meanshift_epsilon;
    out[7] = (float)c.depth_rand_samples, out[8] = c.abs_resize_factor, out[9] = (float)c.exclusive_gpu_context;
    out[10] = c.rg_pose_scaling, out[11] = (float)c.fb_smooth;
}

}  // extern "C"
