// Device-resident VO window: init + EM solve behind py_voldor_wrapper.
//
// Behavioural source: reference voldor/py_export.cpp:5-79 (wrapper, outputs), voldor/voldor.cpp:4-128 (init),
// :130-149 (solve), :164-201 (optimize_cameras incl. truncation), :203-307 (optimize_depth call patterns),
// :309-317 (normalize_world_scale), voldor/geometry.cpp:5-265 (optimize_camera_pose).
//
// What changes against the reference is WHERE data lives, not what is computed: flows are uploaded once and
// shared by the depth step and the collector; rigidness, depth, prior confidences, instance maps, compacted
// instances, hypotheses and the pose pool never leave the GPU; per camera exactly one small result struct
// (pool size, instance count, mean, density, iterations) crosses PCIe.  The reference's cache-visibility
// quirks are preserved: the collector sees the world-scale-normalised depth while the depth step keeps
// optimising its own un-normalised copy, smoothing is applied in place and overwritten by the
// raw E-step (Q18), RNG streams continue across windows (Q1).
#include "../../include/py_export.h"
#include "../../include/voldor_b200.h"
#include "bootstrap.h"
#include "config.h"
#include "context.h"
#include "host_math.h"
#include "residual_model.cuh"
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <mutex>
#include <vector>

namespace vb {

namespace {


struct Camera {  // reference: voldor/utils.h:30-76
    float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    float t[3] = {0, 0, 0};
    float pose_covar[36] = {0};
    float pose_density = 0;
    int pose_sample_count = 0;
    float pose_rigidness_density = 0;
    int last_used_ms_iters = 0, last_used_gu_iters = 0;
};

// ---------------------------------------------------------------------------------------------------
// small kernels
// ---------------------------------------------------------------------------------------------------
__global__ void k_fill(float* p, int pitch, int w, int h, size_t plane, int layers, float v) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    for (int l = 0; l < layers; l++) p[(size_t)l * plane + (size_t)y * pitch + x] = v;
}

// dst = src * s   (host: cv::Mat *= double -> float multiply; reference voldor.cpp:316)
__global__ void k_scale_copy(float* dst, int dpitch, const float* src, int spitch, int w, int h, float s) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    dst[(size_t)y * dpitch + x] = f_mul(src[(size_t)y * spitch + x], s);
}

// dense prior = basefocal / disparity, 0 where the disparity is 0 (reference voldor.cpp:33; OpenCV 3.4 divide)
__global__ void k_disparity_to_depth(float* dst, size_t dpitch_bytes, const float* disp, int w, int h, float bf) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    const float d = disp[(size_t)y * w + x];
    ((float*)((char*)dst + (size_t)y * dpitch_bytes))[x] = d == 0.f ? 0.f : f_div(bf, d);
}

// per-layer sum in double: stage 1 (fixed grid), stage 2 (single block) -> deterministic
constexpr int kSumBlocks = 256;
__global__ void __launch_bounds__(256)
    k_layer_sum_stage1(const float* p, int pitch, int w, int h, size_t plane, double* partial) {
    __shared__ double sh[256];
    const int layer = blockIdx.y;
    const float* src = p + (size_t)layer * plane;
    double acc = 0;
    const int n = w * h;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += kSumBlocks * 256) acc += (double)src[(size_t)(i / w) * pitch + (i % w)];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[layer * kSumBlocks + blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) k_layer_sum_stage2(const double* partial, double* out) {
    __shared__ double sh[256];
    const int layer = blockIdx.x;
    sh[threadIdx.x] = threadIdx.x < kSumBlocks ? partial[layer * kSumBlocks + threadIdx.x] : 0.0;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[layer] = sh[0];
}

// depth_conf = (sum_f rig_f + sum_p conf_p) * (1/(n_f+n_p))   (reference py_export.cpp:66-76)
__global__ void k_depth_conf(float* out, int w, int h, const float* rig, int rpitch, size_t rplane, int nf,
                             const float* conf, int cpitch, size_t cplane, int np, float inv_n) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= w || y >= h) return;
    float acc = 0.f;
    for (int f = 0; f < nf; f++) acc = f_add(acc, rig[(size_t)f * rplane + (size_t)y * rpitch + x]);
    for (int f = 0; f < np; f++) acc = f_add(acc, conf[(size_t)f * cplane + (size_t)y * cpitch + x]);
    out[(size_t)y * w + x] = f_mul(acc, inv_n);
}

// ---------------------------------------------------------------------------------------------------
// the window object (one per call; the device state lives in the execution context it runs on, context.h)
// ---------------------------------------------------------------------------------------------------
struct Window {
    Context& ctx;
    explicit Window(Context& c) : ctx(c), E(c.E), C(c.C), M(c.M) {}
    Config cfg;
    int w = 0, h = 0, n_flows = 0, n_flows_init = 0, n_depth_priors = 0;
    int iters_cur = 0, iters_remain = 0;
    bool has_disparity = false;
    std::vector<Camera> cams;
    float K[9];
    float depth_scale_pending = 1.f;  // host depth == device depth * this (normalize_world_scale)
    bool depth_on_device_valid = false;

    DepthEM& E;
    Collector& C;
    PoseMode& M;
    cudaStream_t s = nullptr;

    // scratch (per context, grow only)
    using Scratch = WindowScratch;
    Scratch& scratch() { return ctx.ws; }

    double t_cameras = 0, t_depth = 0, t_io = 0;

    int ensure_scratch() {
        Scratch& sc = scratch();
        const int P = cfg.n_poses_to_sample;
        if (P > sc.pose_cap) {
            if (sc.rvecs) cudaFree(sc.rvecs), cudaFree(sc.tvecs), cudaFree(sc.pool);
            VB_CUDA(cudaMalloc((void**)&sc.rvecs, (size_t)P * 3 * sizeof(float)));
            VB_CUDA(cudaMalloc((void**)&sc.tvecs, (size_t)P * 3 * sizeof(float)));
            VB_CUDA(cudaMalloc((void**)&sc.pool, (size_t)P * 6 * sizeof(float)));
            sc.pose_cap = P;
        }
        if (!sc.d_used) {
            VB_CUDA(cudaMalloc((void**)&sc.d_used, sizeof(int)));
            VB_CUDA(cudaMalloc((void**)&sc.sum_partial, (size_t)(kMaxFrames + kMaxPriorFrames) * kSumBlocks * sizeof(double)));
            VB_CUDA(cudaMalloc((void**)&sc.sums, (kMaxFrames + kMaxPriorFrames) * sizeof(double)));
            VB_CUDA(cudaMallocHost((void**)&sc.h_sums, (kMaxFrames + kMaxPriorFrames) * sizeof(double)));
            VB_CUDA(cudaMallocHost((void**)&sc.h_counts, 4 * sizeof(int)));
            VB_CUDA(cudaMalloc((void**)&sc.d_cams, sizeof(CamBlock)));
            VB_CUDA(cudaMallocHost((void**)&sc.h_cams, sizeof(CamBlock)));
        }
        const size_t npx = (size_t)w * h;
        if (npx > sc.out_cap) {
            if (sc.d_out) cudaFree(sc.d_out);
            VB_CUDA(cudaMalloc((void**)&sc.d_out, npx * sizeof(float)));
            sc.out_cap = npx;
        }
        return 0;
    }

    void launch2d(dim3& g, dim3& b) const {
        b = dim3(32, 8);
        g = dim3(VB_DIV_CEIL(w, 32), VB_DIV_CEIL(h, 8));
    }

    // ---------------------------------------------------------------------------------------------
    // init (reference voldor.cpp:4-128)
    // ---------------------------------------------------------------------------------------------
    int init(const float* flows_pt, const float* disparity_pt, const float* disparity_pconf_pt,
             const float* depth_priors_pt, const float* depth_prior_poses_pt, const float* depth_prior_pconfs_pt,
             int N, int N_dp_in) {
        if (cfg.resize_factor != 1) {
            printf("voldor_b200: --resize_factor != 1 is not supported (deprecated in the reference; resize in the "
                   "caller as slam_py does)\n");
            return -1;
        }
        n_flows = n_flows_init = N;
        n_depth_priors = N_dp_in + (disparity_pt ? 1 : 0);
        has_disparity = disparity_pt != nullptr;
        if (N > kMaxFrames || n_depth_priors > kMaxPriorFrames || N <= 0) return -1;
        iters_cur = 0;
        iters_remain = cfg.max_iters;
        cams.assign(N, Camera());
        K[0] = cfg.fx, K[1] = 0, K[2] = cfg.cx, K[3] = 0, K[4] = cfg.fy, K[5] = cfg.cy, K[6] = 0, K[7] = 0, K[8] = 1;

        auto t0 = std::chrono::high_resolution_clock::now();
        // device state; flows live once, in the depth step's stack, and the collector aliases them
        E.shared_flows = nullptr;
        if (int e = E.ensure(w, h, N, n_depth_priors)) return e;
        E.overlap_smoothing = true;  // smoothing of iteration k+1's weights overlaps the camera step
        s = E.stream;
        if (int e = C.ensure(w, h, N)) return e;
        if (int e = M.init()) return e;
        if (int e = ensure_scratch()) return e;
        C.depth_own.ensure(w, h, 1, false);
        C.flows = &E.flows;
        C.rig = E.rig.ptr, C.rig_pitch = E.rig.pitch, C.rig_plane = E.rig.layer_elems();
        C.depth = C.depth_own.ptr, C.depth_pitch = C.depth_own.pitch;
        fill_K(C.cam, K);
        E.set_K(K);

        const size_t npx = (size_t)w * h;
        for (int f = 0; f < N; f++) VB_CUDA(E.flows.upload_layer((const float2*)(flows_pt + (size_t)f * npx * 2), f, s));

        dim3 g, b;
        launch2d(g, b);
        // rigidness := 1, prior confidences := 1 (voldor.cpp:89-95)
        k_fill<<<g, b, 0, s>>>(E.rig.ptr, E.rig.pitch, w, h, E.rig.layer_elems(), N, 1.f);
        if (n_depth_priors > 0) {
            const int cp = (int)(E.dp_conf.pitch / sizeof(float));
            k_fill<<<g, b, 0, s>>>(E.dp_conf.ptr, cp, w, h, (size_t)cp * h, n_depth_priors, 1.f);
        }
        int pi = 0;
        std::vector<float> ones;
        if (disparity_pt) {
            Scratch& sc = scratch();
            if (npx > sc.disp_cap) {
                if (sc.d_disp) cudaFree(sc.d_disp);
                VB_CUDA(cudaMalloc((void**)&sc.d_disp, npx * sizeof(float)));
                sc.disp_cap = npx;
            }
            VB_CUDA(cudaMemcpyAsync(sc.d_disp, disparity_pt, npx * sizeof(float), cudaMemcpyDefault, s));
            k_disparity_to_depth<<<g, b, 0, s>>>(E.dp.layer(0), E.dp.pitch, sc.d_disp, w, h, cfg.basefocal);
            if (disparity_pconf_pt) {
                VB_CUDA(E.dp_pconf.upload_layer(disparity_pconf_pt, 0, s));
            } else {
                const int pp = (int)(E.dp_pconf.pitch / sizeof(float));
                k_fill<<<g, b, 0, s>>>(E.dp_pconf.layer(0), pp, w, h, (size_t)pp * h, 1, 1.f);
            }
            const float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, z[3] = {0, 0, 0};
            E.set_prior_pose(0, I, z);
            pi = 1;
        }
        for (int i = 0; i < N_dp_in; i++, pi++) {
            VB_CUDA(E.dp.upload_layer(depth_priors_pt + (size_t)i * npx, pi, s));
            if (depth_prior_pconfs_pt) {
                VB_CUDA(E.dp_pconf.upload_layer(depth_prior_pconfs_pt + (size_t)i * npx, pi, s));
            } else {
                const int pp = (int)(E.dp_pconf.pitch / sizeof(float));
                k_fill<<<g, b, 0, s>>>(E.dp_pconf.layer(pi), pp, w, h, (size_t)pp * h, 1, 1.f);
            }
            float R[9];
            hm::rvec_to_matrix(depth_prior_poses_pt + i * 6, R);
            E.set_prior_pose(pi, R, depth_prior_poses_pt + i * 6 + 3);
        }
        // depth init (voldor.cpp:111-121)
        if (n_depth_priors > 0) {
            VB_CUDA(cudaMemcpy2DAsync(E.depth.ptr, (size_t)E.depth.pitch * sizeof(float), E.dp.layer(0), E.dp.pitch,
                                      (size_t)w * sizeof(float), h, cudaMemcpyDeviceToDevice, s));
            if (!disparity_pt) {
                if (int e = optimize_depth(/*only_prior=*/true, false)) return e;
            }
        } else {
            k_fill<<<g, b, 0, s>>>(E.depth.ptr, E.depth.pitch, w, h, E.depth.layer_elems(), 1, 1.f);
        }
        VB_RETURN_IF_CUDA_ERROR();
        depth_scale_pending = 1.f;
        t_io += std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        return 0;
    }

    DepthHyper hyper() const {
        DepthHyper hp;
        hp.abs_resize_factor = cfg.abs_resize_factor, hp.basefocal = cfg.basefocal;
        hp.n_rand_samples = cfg.depth_rand_samples, hp.global_prop_step = cfg.depth_global_prop_step;
        hp.local_prop_width = cfg.depth_local_prop_width;
        hp.lambda = cfg.lambda, hp.omega = cfg.omega;
        hp.disp_delta = has_disparity ? cfg.disp_delta : -1, hp.delta = cfg.delta;
        hp.fb_smooth = cfg.fb_smooth != 0, hp.s0_ems_prob = cfg.fb_emm, hp.no_change_prob = cfg.fb_no_change_prob;
        hp.range_factor = cfg.depth_range_factor;
        return hp;
    }

    // reference voldor.cpp:203-307
    int optimize_depth(bool only_prior, bool rigidness_only) {
        if (n_flows == 0 && n_depth_priors == 0) return 0;
        const int N = only_prior ? 0 : n_flows;
        for (int f = 0; f < N; f++) E.set_pose(f, cams[f].R, cams[f].t);
        if (!cfg.exclusive_gpu_context && depth_scale_pending != 1.f) {
            // non-exclusive mode re-uploads the (normalised) host depth on every call (voldor.cpp:250-270)
            dim3 g, b;
            launch2d(g, b);
            k_scale_copy<<<g, b, 0, s>>>(E.depth.ptr, E.depth.pitch, E.depth.ptr, E.depth.pitch, w, h, depth_scale_pending);
            depth_scale_pending = 1.f;
        }
        return E.run(N, n_depth_priors, hyper(), rigidness_only);
    }

    // reference geometry.cpp:5-265; returns 1 on success
    int optimize_camera_pose(int i, bool successive_pose, bool rg_refine, bool refresh_depth) {
        Scratch& sc = scratch();
        Camera& cam = cams[i];
        for (int f = 0; f < n_flows; f++) {
            memcpy(C.cam.R[f], cams[f].R, 9 * sizeof(float));
            memcpy(C.cam.t[f], cams[f].t, 3 * sizeof(float));
        }
        if (refresh_depth) {
            // the collector works on the host's view of the depth map = device depth * pending world scale
            dim3 g, b;
            launch2d(g, b);
            k_scale_copy<<<g, b, 0, s>>>(C.depth_own.ptr, C.depth_own.pitch, E.depth.ptr, E.depth.pitch, w, h,
                                         depth_scale_pending);
        }
        // everything below is enqueued on the depth step's stream so it is ordered after the E-step
        C.stream = s;
        CollectParams P;
        P.active_idx = i, P.rigidness_thresh = cfg.rigidness_threshold;
        P.rigidness_sum_thresh = cfg.rigidness_sum_threshold;
        P.sample_min_depth = cfg.pose_sample_min_depth, P.sample_max_depth = cfg.pose_sample_max_depth;
        P.max_trace_on_flow = cfg.max_trace_on_flow;
        if (int e = C.collect(n_flows, P, true)) return e < 0 ? e : -e;
        if (solve_batch_p3p_device(C.p3c, C.p2c, C.d_count, 0, K[0], K[4], K[2], K[5], sc.rvecs, sc.tvecs,
                                   cfg.n_poses_to_sample, !cfg.lambdatwist, s))
            return -1;
        // successive poses: the finite filter of the hypotheses is fused into the mean-shift launch
        if (!successive_pose)
            if (filter_pose_pool(sc.rvecs, sc.tvecs, cfg.n_poses_to_sample, cfg.meanshift_rvec_scale, sc.pool, sc.d_used, s))
                return -1;

        float pose_opm[6];
        hm::matrix_to_rvec(cam.R, pose_opm);
        pose_opm[3] = cam.t[0], pose_opm[4] = cam.t[1], pose_opm[5] = cam.t[2];
        for (int d = 0; d < 3; d++) pose_opm[d] *= cfg.meanshift_rvec_scale;

        int n_points = -1, pool_used = -1;
        M.stream = s;
        if (!successive_pose) {
            // the start-sample trials need the pool size on the host anyway
            cudaMemcpyAsync(&sc.h_counts[0], C.d_count, sizeof(int), cudaMemcpyDeviceToHost, s);
            cudaMemcpyAsync(&sc.h_counts[1], sc.d_used, sizeof(int), cudaMemcpyDeviceToHost, s);
            cudaStreamSynchronize(s);
            n_points = sc.h_counts[0], pool_used = sc.h_counts[1];
            if (n_points < 4) return 0;
            if (pool_used == 0) return 0;
            if (M.meanshift(sc.pool, nullptr, nullptr, pool_used, 6, cfg.meanshift_kernel_var, pose_opm,
                            &cam.pose_density, &cam.last_used_ms_iters, false, cfg.meanshift_epsilon,
                            cfg.meanshift_max_iters, cfg.meanshift_max_init_trials, cfg.meanshift_good_init_confidence))
                return -1;
        } else {
            float density = cam.pose_density;
            int ms_iters = cam.last_used_ms_iters;
            float mean_io[6];
            memcpy(mean_io, pose_opm, sizeof(mean_io));
            if (M.meanshift_from_hypotheses(sc.rvecs, sc.tvecs, cfg.n_poses_to_sample, cfg.meanshift_rvec_scale, sc.pool,
                                            sc.d_used, 6, cfg.meanshift_kernel_var, mean_io, &density, &ms_iters,
                                            cfg.meanshift_epsilon, cfg.meanshift_max_iters, C.d_count))
                return -1;
            n_points = M.h_result->aux_count;
            pool_used = M.h_result->n;
            if (n_points < 4) return 0;
            if (pool_used == 0) return 0;
            memcpy(pose_opm, mean_io, sizeof(mean_io));
            cam.pose_density = density, cam.last_used_ms_iters = ms_iters;
        }
        cam.pose_sample_count = pool_used;

        if (rg_refine) {
            // reference geometry.cpp:201-246
            const float sc2 = cfg.rg_pose_scaling * cfg.rg_pose_scaling;
            for (int k = 0; k < 36; k++) cam.pose_covar[k] = 0;
            for (int d = 0; d < 6; d++) cam.pose_covar[d * 6 + d] = cfg.meanshift_kernel_var;
            for (int k = 0; k < 36; k++) cam.pose_covar[k] *= sc2;
            for (int d = 0; d < 6; d++) pose_opm[d] *= cfg.rg_pose_scaling;
            const int ret = M.fit_robust_gaussian(sc.pool, pool_used, 6, cfg.rg_pose_scaling, pose_opm, cam.pose_covar,
                                                  cfg.rg_trunc_sigma, cfg.rg_covar_reg_lambda, &cam.pose_density,
                                                  &cam.last_used_gu_iters, cfg.rg_epsilon, cfg.rg_max_iters);
            if (ret == 0) {
                const float inv = (float)(1. / (double)sc2);  // cv::Mat /= s multiplies by 1/s
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
            if (!std::isfinite(pose_opm[d])) return 0;  // cv::checkRange
        hm::rvec_to_matrix(pose_opm, cam.R);
        cam.t[0] = pose_opm[3], cam.t[1] = pose_opm[4], cam.t[2] = pose_opm[5];
        return 1;
    }

    int fetch_rigidness_densities() {
        Scratch& sc = scratch();
        k_layer_sum_stage1<<<dim3(kSumBlocks, n_flows), 256, 0, s>>>(E.rig.ptr, E.rig.pitch, w, h, E.rig.layer_elems(),
                                                                     sc.sum_partial);
        k_layer_sum_stage2<<<n_flows, 256, 0, s>>>(sc.sum_partial, sc.sums);
        VB_CUDA(cudaMemcpyAsync(sc.h_sums, sc.sums, n_flows * sizeof(double), cudaMemcpyDeviceToHost, s));
        VB_CUDA(cudaStreamSynchronize(s));
        for (int i = 0; i < n_flows; i++) cams[i].pose_rigidness_density = (float)sc.h_sums[i] / (float)(w * h);
        return 0;
    }

    // reference voldor.cpp:164-201
    // The camera loop of one EM iteration without a host round trip per camera (reference voldor.cpp:170-195 +
    // geometry.cpp:5-265, successive poses, no robust refinement): the kernels of all cameras are enqueued back to
    // back; the mean-shift launch of camera i derives its R,t on the device (PoseTail) and stores them in the
    // device-resident camera block the next camera's collection kernel reads.  One copy + one synchronisation per
    // iteration; the host then replays the reference's sequential bookkeeping (failure / truncation) on the
    // results, discarding what was computed past a truncation point.
    int optimize_cameras_pipelined(bool allow_trunc) {
        Scratch& sc = scratch();
        for (int f = 0; f < n_flows; f++) {
            memcpy(C.cam.R[f], cams[f].R, 9 * sizeof(float));
            memcpy(C.cam.t[f], cams[f].t, 3 * sizeof(float));
        }
        *sc.h_cams = C.cam;
        VB_CUDA(cudaMemcpyAsync(sc.d_cams, sc.h_cams, sizeof(CamBlock), cudaMemcpyHostToDevice, s));
        dim3 g, b;
        launch2d(g, b);
        k_scale_copy<<<g, b, 0, s>>>(C.depth_own.ptr, C.depth_own.pitch, E.depth.ptr, E.depth.pitch, w, h,
                                     depth_scale_pending);
        C.stream = s, M.stream = s;
        C.d_cam = sc.d_cams;
        int planned = n_flows;  // cameras to attempt; a camera skipped by the density gate truncates right there
        for (int i = 0; i < n_flows; i++) {
            if (allow_trunc && !(cams[i].pose_rigidness_density > cfg.trunc_rigidness_density)) {
                planned = i;
                break;
            }
            CollectParams P;
            P.active_idx = i, P.rigidness_thresh = cfg.rigidness_threshold;
            P.rigidness_sum_thresh = cfg.rigidness_sum_threshold;
            P.sample_min_depth = cfg.pose_sample_min_depth, P.sample_max_depth = cfg.pose_sample_max_depth;
            P.max_trace_on_flow = cfg.max_trace_on_flow;
            int e = C.collect(n_flows, P, true);
            if (!e)
                e = solve_batch_p3p_device(C.p3c, C.p2c, C.d_count, 0, K[0], K[4], K[2], K[5], sc.rvecs, sc.tvecs,
                                           cfg.n_poses_to_sample, !cfg.lambdatwist, s);
            float pose_opm[6];
            hm::matrix_to_rvec(cams[i].R, pose_opm);
            pose_opm[3] = cams[i].t[0], pose_opm[4] = cams[i].t[1], pose_opm[5] = cams[i].t[2];
            for (int d = 0; d < 3; d++) pose_opm[d] *= cfg.meanshift_rvec_scale;
            PoseTail tail;
            tail.d_cams = sc.d_cams, tail.cam_index = i;
            tail.inv_rvec_scale = (float)(1. / (double)cfg.meanshift_rvec_scale);
            if (!e)
                e = M.enqueue_from_hypotheses(i, sc.rvecs, sc.tvecs, cfg.n_poses_to_sample, cfg.meanshift_rvec_scale,
                                              sc.pool, sc.d_used, 6, cfg.meanshift_kernel_var, pose_opm,
                                              cfg.meanshift_epsilon, cfg.meanshift_max_iters, C.d_count, tail);
            if (e) {
                C.d_cam = nullptr;
                return -1;
            }
        }
        C.d_cam = nullptr;
        if (M.fetch_results(planned)) return -1;
        for (int i = 0; i < n_flows; i++) {
            int ok = 0;
            Camera& cam = cams[i];
            if (i < planned) {
                const MeanshiftResult& r = M.h_result[i];
                if (r.aux_count >= 4 && r.n > 0) {
                    if (r.used_iters > 0) cam.pose_density = r.confidence;
                    cam.last_used_ms_iters = r.used_iters;
                    cam.pose_sample_count = r.n;
                    if (r.ok) {
                        memcpy(cam.R, r.R, sizeof(cam.R));
                        memcpy(cam.t, r.t, sizeof(cam.t));
                        ok = 1;
                    }
                }
            }
            if (!ok || (allow_trunc && cam.pose_density < cfg.trunc_sample_density)) {
                iters_remain = std::max(iters_remain, cfg.min_iters_after_trunc);
                n_flows = i;
                break;
            }
        }
        return 0;
    }

    int optimize_cameras() {
        const bool allow_trunc = iters_cur > cfg.no_trunc_iters;
        if (allow_trunc)
            if (int e = fetch_rigidness_densities()) return e;
        // pipelined loop when every camera continues from its previous pose and no robust refinement is due
        static const bool pipeline_off = getenv("VB_NO_CAMERA_PIPELINE") != nullptr;
        bool pipelined = !pipeline_off && n_flows > 0 &&
                         !(cfg.rg_refine && (!cfg.rg_refine_last_only || iters_remain == 0));
        for (int i = 0; i < n_flows && pipelined; i++) pipelined = cams[i].pose_sample_count != 0;
        if (pipelined) return optimize_cameras_pipelined(allow_trunc);
        for (int i = 0; i < n_flows; i++) {
            int ok = 0;
            if (!allow_trunc || cams[i].pose_rigidness_density > cfg.trunc_rigidness_density) {
                ok = optimize_camera_pose(i, cams[i].pose_sample_count != 0,
                                          cfg.rg_refine && (!cfg.rg_refine_last_only || iters_remain == 0), i == 0);
                if (ok < 0) return ok;
            }
            if (!ok || (allow_trunc && cams[i].pose_density < cfg.trunc_sample_density)) {
                iters_remain = std::max(iters_remain, cfg.min_iters_after_trunc);
                n_flows = i;
                break;
            }
        }
        return 0;
    }

    // reference voldor.cpp:309-317
    void normalize_world_scale() {
        float world_scale = 0;
        for (int i = 0; i < n_flows; i++) world_scale = (float)((double)world_scale + hm::norm3(cams[i].t));
        const float sc = n_flows / world_scale;
        for (int i = 0; i < n_flows; i++)
            for (int d = 0; d < 3; d++) cams[i].t[d] *= sc;
        depth_scale_pending = sc;  // host depth = device depth * sc (Q2)
    }

    int bootstrap(const float* flows_pt) {
        const BootstrapOverride& g_bootstrap = ctx.boot;
        if (g_bootstrap.valid && g_bootstrap.w == w && g_bootstrap.h == h) {
            memcpy(cams[0].R, g_bootstrap.R, sizeof(g_bootstrap.R));
            memcpy(cams[0].t, g_bootstrap.t, sizeof(g_bootstrap.t));
            VB_CUDA(E.depth.upload_layer(g_bootstrap.depth.data(), 0, s));
            VB_CUDA(cudaStreamSynchronize(s));
            return 0;
        }
        // essential-matrix bootstrap + closed-form depth on the host (reference geometry.cpp:267-332)
        // flows_pt may be a host or a device pointer; the estimator runs on the host
        std::vector<float> depth((size_t)w * h), flow0((size_t)w * h * 2);
        VB_CUDA(cudaMemcpy(flow0.data(), flows_pt, flow0.size() * sizeof(float), cudaMemcpyDefault));
        if (!boot::bootstrap_from_flow(flow0.data(), w, h, K, cams[0].R, cams[0].t, depth.data())) {
            fprintf(stderr, "voldor_b200: monocular bootstrap failed (degenerate flow)\n");
            return -1;
        }
        VB_CUDA(E.depth.upload_layer(depth.data(), 0, s));
        VB_CUDA(cudaStreamSynchronize(s));
        return 0;
    }

    // reference voldor.cpp:130-149
    int solve(const float* flows_pt) {
        using clk = std::chrono::high_resolution_clock;
        if (n_depth_priors == 0)
            if (int e = bootstrap(flows_pt)) return e;
        while (iters_remain > 0 && n_flows > 0) {
            iters_cur++;
            iters_remain--;
            auto t0 = clk::now();
            if (int e = optimize_cameras()) return e;
            auto t1 = clk::now();
            if (int e = optimize_depth(false, !cfg.optimize_depth)) return e;
            if (cfg.norm_world_scale && n_depth_priors == 0) normalize_world_scale();
            auto t2 = clk::now();
            t_cameras += std::chrono::duration<double, std::milli>(t1 - t0).count();
            t_depth += std::chrono::duration<double, std::milli>(t2 - t1).count();
        }
        return 0;
    }

    // reference py_export.cpp:56-76
    int outputs(int* n_registered, float* poses_pt, float* poses_covar_pt, float* depth_pt, float* depth_conf_pt) {
        Scratch& sc = scratch();
        *n_registered = n_flows;
        for (int i = 0; i < n_flows; i++) {
            if (poses_pt) {
                hm::matrix_to_rvec(cams[i].R, poses_pt + i * 6);
                memcpy(poses_pt + i * 6 + 3, cams[i].t, 3 * sizeof(float));
            }
            if (poses_covar_pt) memcpy(poses_covar_pt + i * 36, cams[i].pose_covar, 36 * sizeof(float));
        }
        dim3 g, b;
        launch2d(g, b);
        const size_t npx = (size_t)w * h;
        // A window that lost every camera ends with 0/0 factors (world scale, 1/(n_flows+n_priors)).  The reference
        // applies them on the host, where x86 produces the default NaN 0xFFC00000 while the GPU would produce
        // 0x7FFFFFFF; those degenerate maps are finished on the host so that even the failure outputs match.
        if (depth_pt) {
            if (std::isfinite(depth_scale_pending)) {
                k_scale_copy<<<g, b, 0, s>>>(sc.d_out, w, E.depth.ptr, E.depth.pitch, w, h, depth_scale_pending);
                VB_CUDA(cudaMemcpyAsync(depth_pt, sc.d_out, npx * sizeof(float), cudaMemcpyDefault, s));
                VB_CUDA(cudaStreamSynchronize(s));
            } else {
                std::vector<float> raw(npx);
                VB_CUDA(cudaMemcpy2DAsync(raw.data(), (size_t)w * sizeof(float), E.depth.ptr,
                                          (size_t)E.depth.pitch * sizeof(float), (size_t)w * sizeof(float), h,
                                          cudaMemcpyDeviceToHost, s));
                VB_CUDA(cudaStreamSynchronize(s));
                volatile float scale = depth_scale_pending;
                for (size_t k = 0; k < npx; k++) raw[k] = raw[k] * scale;
                VB_CUDA(cudaMemcpy(depth_pt, raw.data(), npx * sizeof(float), cudaMemcpyDefault));
            }
        }
        if (depth_conf_pt) {
            const float inv_n = (float)(1. / (double)(float)(n_flows + n_depth_priors));
            if (n_flows + n_depth_priors > 0) {
                const int cp = (int)(E.dp_conf.pitch / sizeof(float));
                k_depth_conf<<<g, b, 0, s>>>(sc.d_out, w, h, E.rig.ptr, E.rig.pitch, E.rig.layer_elems(), n_flows,
                                             E.dp_conf.ptr, cp, (size_t)cp * h, n_depth_priors, inv_n);
                VB_CUDA(cudaMemcpyAsync(depth_conf_pt, sc.d_out, npx * sizeof(float), cudaMemcpyDefault, s));
                VB_CUDA(cudaStreamSynchronize(s));
            } else {
                volatile float zero = 0.f, f = inv_n;
                std::vector<float> conf(npx, zero * f);
                VB_CUDA(cudaMemcpy(depth_conf_pt, conf.data(), npx * sizeof(float), cudaMemcpyDefault));
            }
        }
        VB_RETURN_IF_CUDA_ERROR();
        return 0;
    }
};

int run_window(const float* flows_pt, const float* disparity_pt, const float* disparity_pconf_pt,
               const float* depth_priors_pt, const float* depth_prior_poses_pt, const float* depth_prior_pconfs_pt,
               float fx, float fy, float cx, float cy, float basefocal, int N, int N_dp, int w, int h,
               const char* config_pt, int* n_registered, float* poses_pt, float* poses_covar_pt, float* depth_pt,
               float* depth_conf_pt, int* iters_run, float* stats) {
    enter_device();
    Context& context = current_context();
    std::lock_guard<std::recursive_mutex> lock(context.mutex);
    auto t0 = std::chrono::high_resolution_clock::now();
    Window W(context);
    W.cfg.fx = fx, W.cfg.cx = cx, W.cfg.fy = fy, W.cfg.cy = cy, W.cfg.basefocal = basefocal;
    W.cfg.read(config_pt);
    if (W.cfg.cpu_p3p) {
        printf("voldor_b200: --cpu_p3p 1 is not available (no CPU fallback in this build)\n");
        *n_registered = 0;
        return 0;
    }
    W.w = w, W.h = h;
    *n_registered = 0;
    int rc = W.init(flows_pt, disparity_pt, disparity_pconf_pt, depth_priors_pt, depth_prior_poses_pt,
                    depth_prior_pconfs_pt, N, N_dp);
    if (rc == 0) rc = W.solve(flows_pt);
    if (rc == 0) {
        auto t1 = std::chrono::high_resolution_clock::now();
        rc = W.outputs(n_registered, poses_pt, poses_covar_pt, depth_pt, depth_conf_pt);
        W.t_io += std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t1).count();
    }
    if (rc != 0) {
        printf("voldor_b200: window failed (code %d)\n", rc);
        *n_registered = 0;
    }
    if (iters_run) *iters_run = W.iters_cur;
    if (stats) {
        stats[0] = (float)std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
        stats[1] = (float)W.t_cameras, stats[2] = (float)W.t_depth, stats[3] = (float)W.t_io;
    }
    return 0;  // the reference wrapper always returns 0; failure is n_registered == 0 (py_export.cpp:78)
}

}  // namespace
}  // namespace vb

int py_voldor_wrapper(const float* flows, const float* disparity, const float* disparity_pconf,
                      const float* depth_priors, const float* depth_prior_poses, const float* depth_prior_pconfs,
                      const float fx, const float fy, const float cx, const float cy, const float basefocal, const int N,
                      const int N_dp, const int w, const int h, const char* config, int& n_registered, float* poses,
                      float* poses_covar, float* depth, float* depth_conf) {
    return vb::run_window(flows, disparity, disparity_pconf, depth_priors, depth_prior_poses, depth_prior_pconfs, fx, fy,
                          cx, cy, basefocal, N, N_dp, w, h, config, &n_registered, poses, poses_covar, depth, depth_conf,
                          nullptr, nullptr);
}

extern "C" {

VB_EXPORT int vb_py_voldor_wrapper(const float* flows, const float* disparity, const float* disparity_pconf,
                                   const float* depth_priors, const float* depth_prior_poses,
                                   const float* depth_prior_pconfs, float fx, float fy, float cx, float cy,
                                   float basefocal, int N, int N_dp, int w, int h, const char* config, int* n_registered,
                                   float* poses, float* poses_covar, float* depth, float* depth_conf) {
    return vb::run_window(flows, disparity, disparity_pconf, depth_priors, depth_prior_poses, depth_prior_pconfs, fx, fy,
                          cx, cy, basefocal, N, N_dp, w, h, config, n_registered, poses, poses_covar, depth, depth_conf,
                          nullptr, nullptr);
}

VB_EXPORT int vb_py_voldor_wrapper_ex(const float* flows, const float* disparity, const float* disparity_pconf,
                                      const float* depth_priors, const float* depth_prior_poses,
                                      const float* depth_prior_pconfs, float fx, float fy, float cx, float cy,
                                      float basefocal, int N, int N_dp, int w, int h, const char* config,
                                      int* n_registered, float* poses, float* poses_covar, float* depth,
                                      float* depth_conf, int* iters_run, float* stats) {
    return vb::run_window(flows, disparity, disparity_pconf, depth_priors, depth_prior_poses, depth_prior_pconfs, fx, fy,
                          cx, cy, basefocal, N, N_dp, w, h, config, n_registered, poses, poses_covar, depth, depth_conf,
                          iters_run, stats);
}

// Test hook: the flag parser of this library (csrc/config.h) on a flag string, every field dumped in the order of
// oracle/ref_shim/ref_config_probe.cpp (which dumps the reference's own Config for the same string).
VB_EXPORT int vb_debug_config_dump(const char* flags, float fx, float fy, float cx, float cy, float basefocal,
                                   double* out) {
    vb::Config cfg;
    cfg.fx = fx, cfg.cx = cx, cfg.fy = fy, cfg.cy = cy, cfg.basefocal = basefocal;
    cfg.read(flags);
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

VB_EXPORT int vb_bootstrap_from_flow(const float* flow, int w, int h, const float* K9, float* R9, float* t3, float* depth) {
    return vb::boot::bootstrap_from_flow(flow, w, h, K9, R9, t3, depth) ? 0 : 1;
}

VB_EXPORT int vb_set_bootstrap_override(int valid, const float* R9, const float* t3, const float* depth, int w, int h) {
    vb::Context& cx = vb::current_context();  // the override belongs to the calling thread's execution context
    std::lock_guard<std::recursive_mutex> lock(cx.mutex);
    vb::BootstrapOverride& b = cx.boot;
    b.valid = valid != 0;
    if (!valid) return 0;
    memcpy(b.R, R9, 9 * sizeof(float));
    memcpy(b.t, t3, 3 * sizeof(float));
    b.depth.assign(depth, depth + (size_t)w * h);
    b.w = w, b.h = h;
    return 0;
}

}  // extern "C"
