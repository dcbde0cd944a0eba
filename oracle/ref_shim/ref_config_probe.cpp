// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// The reference's own `struct Config` (voldor/config.h:4-253, taken from the reference tree by -I at build time,
// unmodified) exposed to ctypes: construct, set the intrinsics and parse a flag string exactly like
// voldor/py_export.cpp:15-25 does, then dump every field.  tests/test_cpu_config_pin.py compares the dump with the
// product's flag parser (csrc/config.h through vb_debug_config_dump) — defaults, the switch fall-through of str_to_arg
// (config.h:85-99: every numeric flag ends as stod cast to the field's type), value-less switches.
#include <iterator>
#include <sstream>
#include "config.h"

extern "C" int ref_config_dump(const char* flags, float fx, float fy, float cx, float cy, float basefocal, double* out) {
    Config cfg;
    std::istringstream iss(flags);
    std::vector<std::string> cfg_strs(std::istream_iterator<std::string>{iss}, std::istream_iterator<std::string>());
    cfg.fx = fx, cfg.cx = cx, cfg.fy = fy, cfg.cy = cy, cfg.basefocal = basefocal;
    cfg.read_config(cfg_strs);
    int k = 0;
#define F(name) out[k++] = (double)cfg.name
    F(omega); F(disp_delta); F(delta); F(basefocal);
    F(rg_refine); F(rg_refine_last_only); F(rg_trunc_sigma); F(rg_covar_reg_lambda); F(rg_pose_scaling); F(rg_max_iters); F(rg_epsilon);
    F(resize_factor); F(abs_resize_factor); F(fx); F(fy); F(cx); F(cy); F(exclusive_gpu_context);
    F(debug); F(silent); F(save_everything); F(viz_img_per_row); F(viz_depth_scale);
    F(lambda); F(meanshift_kernel_var); F(meanshift_rvec_scale); F(norm_world_scale);
    F(cpu_p3p); F(lambdatwist); F(n_poses_to_sample); F(pose_sample_min_depth); F(pose_sample_max_depth); F(max_trace_on_flow);
    F(rigidness_threshold); F(rigidness_sum_threshold);
    F(trunc_rigidness_density); F(trunc_sample_density); F(no_trunc_iters); F(max_iters); F(min_iters_after_trunc);
    F(fb_smooth); F(fb_emm); F(fb_no_change_prob);
    F(optimize_depth); F(depth_rand_samples); F(depth_global_prop_step); F(depth_local_prop_width); F(depth_range_factor);
    F(meanshift_max_iters); F(meanshift_max_init_trials); F(meanshift_good_init_confidence); F(meanshift_epsilon);
    F(kitti_estimate_ground); F(kitti_ground_holo_width); F(kitti_ground_roi); F(kitti_ground_meanshift_kernel_var);
#undef F
    return k;
}
