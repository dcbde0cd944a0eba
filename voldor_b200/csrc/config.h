// Window configuration: defaults and the "--flag value" string grammar of the reference.
//
// Behavioural source: reference voldor/config.h:4-82 (fields and defaults), :85-99 (numeric parsing: every
// numeric flag ends up as (T)stod(value) because of the switch fall-through), :110-253 (flag
// names; --debug/--silent/--save_everything are value-less switches; unknown flag or missing value -> message
// and exit(1)).  Transported as one C string through py_voldor_wrapper (voldor/py_export.cpp:15-25).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <sstream>
#include <string>
#include <vector>

namespace vb {

struct Config {
    // depth prior related
    float omega = 0.15f, disp_delta = 1.f, delta = 0.5f, basefocal = 0;
    // robust gaussian fit related
    int rg_refine = true, rg_refine_last_only = true;
    float rg_trunc_sigma = 3.f, rg_covar_reg_lambda = 0.001f, rg_pose_scaling = 100.f;
    int rg_max_iters = 100;
    float rg_epsilon = 1e-5f;
    // input params
    float resize_factor = 1.0f, abs_resize_factor = 1.0f;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    int exclusive_gpu_context = true;
    // debug related
    bool debug = false, silent = false, save_everything = false;
    int viz_img_per_row = 2;
    float viz_depth_scale = 5;
    // hyper-params
    float lambda = 0.15f, meanshift_kernel_var = 0.1f, meanshift_rvec_scale = 25.0f;
    int norm_world_scale = true;
    // pose sampling related
    int cpu_p3p = false, lambdatwist = true, n_poses_to_sample = 8192;
    float pose_sample_min_depth = 0.1f, pose_sample_max_depth = 1000.0f;
    int max_trace_on_flow = 3;
    float rigidness_threshold = 0.5f, rigidness_sum_threshold = 1.f;
    // truncation related
    float trunc_rigidness_density = 0.05f, trunc_sample_density = 0.001f, no_trunc_iters = 2;
    int max_iters = 5, min_iters_after_trunc = 3;
    // fb smooth related
    int fb_smooth = true;
    float fb_emm = 0.5f, fb_no_change_prob = 0.9f;
    // depth update related
    int optimize_depth = true, depth_rand_samples = 10, depth_global_prop_step = 8, depth_local_prop_width = 32;
    float depth_range_factor = 1.f;
    // meanshift related
    int meanshift_max_iters = 100, meanshift_max_init_trials = 20;
    float meanshift_good_init_confidence = 0.5f, meanshift_epsilon = 1e-5;
    // KITTI ground (deprecated in the reference; parsed, not acted on)
    int kitti_estimate_ground = false, kitti_ground_holo_width = 5;
    float kitti_ground_roi = 0.4f, kitti_ground_meanshift_kernel_var = 0.01f;

    template <typename T>
    static void num(const std::string& s, T& arg) {
        arg = (T)std::stod(s);
    }

    void read(const char* config_str) {
        std::istringstream iss(config_str ? config_str : "");
        std::vector<std::string> v;
        for (std::string tok; iss >> tok;) v.push_back(tok);
        auto next = [&](size_t& i) -> const std::string& {
            if (i + 1 < v.size()) return v[++i];
            printf("Config array index out of bound.\n");
            exit(1);
        };
        for (size_t i = 0; i < v.size(); i++) {
            const std::string& k = v[i];
#define VB_FLAG(name) else if (k == "--" #name) num(next(i), this->name)
            if (0) {
            }
            VB_FLAG(basefocal);
            VB_FLAG(omega);
            VB_FLAG(disp_delta);
            VB_FLAG(delta);
            VB_FLAG(rg_refine);
            VB_FLAG(rg_refine_last_only);
            VB_FLAG(rg_trunc_sigma);
            VB_FLAG(rg_covar_reg_lambda);
            VB_FLAG(rg_epsilon);
            VB_FLAG(rg_max_iters);
            VB_FLAG(rg_pose_scaling);
            VB_FLAG(resize_factor);
            VB_FLAG(abs_resize_factor);
            VB_FLAG(fx);
            VB_FLAG(fy);
            VB_FLAG(cx);
            VB_FLAG(cy);
            else if (k == "--debug") debug = true;
            else if (k == "--silent") silent = true;
            else if (k == "--save_everything") save_everything = true;
            VB_FLAG(viz_img_per_row);
            VB_FLAG(viz_depth_scale);
            VB_FLAG(exclusive_gpu_context);
            VB_FLAG(lambda);
            VB_FLAG(meanshift_kernel_var);
            VB_FLAG(meanshift_rvec_scale);
            VB_FLAG(norm_world_scale);
            VB_FLAG(cpu_p3p);
            VB_FLAG(lambdatwist);
            VB_FLAG(max_trace_on_flow);
            VB_FLAG(n_poses_to_sample);
            VB_FLAG(pose_sample_min_depth);
            VB_FLAG(pose_sample_max_depth);
            VB_FLAG(rigidness_threshold);
            VB_FLAG(rigidness_sum_threshold);
            VB_FLAG(trunc_rigidness_density);
            VB_FLAG(trunc_sample_density);
            VB_FLAG(max_iters);
            VB_FLAG(no_trunc_iters);
            VB_FLAG(min_iters_after_trunc);
            VB_FLAG(fb_smooth);
            VB_FLAG(fb_emm);
            VB_FLAG(fb_no_change_prob);
            VB_FLAG(optimize_depth);
            VB_FLAG(depth_rand_samples);
            VB_FLAG(depth_global_prop_step);
            VB_FLAG(depth_local_prop_width);
            VB_FLAG(depth_range_factor);
            VB_FLAG(meanshift_max_iters);
            VB_FLAG(meanshift_max_init_trials);
            VB_FLAG(meanshift_good_init_confidence);
            VB_FLAG(meanshift_epsilon);
            VB_FLAG(kitti_estimate_ground);
            VB_FLAG(kitti_ground_holo_width);
            VB_FLAG(kitti_ground_roi);
            VB_FLAG(kitti_ground_meanshift_kernel_var);
            else {
                printf("Invalid input config : %s\n", k.c_str());
                exit(1);
            }
#undef VB_FLAG
        }
    }
};

}  // namespace vb
