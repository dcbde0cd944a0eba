// TEST INFRASTRUCTURE.  Device build of the product's quad-lane minimal solvers (voldor_b200/csrc/p3p_*_quad.cuh) with
// the undecided contraction sites switchable at run time, so that tests/test_gpu_p3p_sites.py can try every assignment
// against the reference kernels on the GPU (the AP3P solver calls cbrtf/atan2f/powf/cosf, whose host versions differ
// from libdevice in the last bits, so its sites cannot be settled on the CPU like lambda-twist's).
// One thread per hypothesis runs the four lanes of a quad one after the other — same arithmetic as the product kernel.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -o tests/_build/libp3p_probe.so tests/p3p_device_probe.cu
#include <cuda_runtime.h>
#include <curand_kernel.h>

__constant__ unsigned c_twist_sites, c_ap3p_sites;
#define VBQ_SITE_SEARCH_DEVICE
#include "../voldor_b200/csrc/p3p_twist_quad.cuh"
#include "../voldor_b200/csrc/p3p_ap3p_quad.cuh"
#include "../voldor_b200/csrc/rotation.cuh"

using namespace vb::quad;

__global__ void k_probe(int solver, const float* p2s, const float* p3s, int n_pts, int n_poses, float fx, float fy, float cx,
                        float cy, float* rvecs, float* tvecs) {
    const int h = blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= n_poses) return;
    curandStateXORWOW_t st;
    curand_init(233ULL, (unsigned long long)h, 0, &st);
    float uv[8];
    Vec3f X[4];
    for (int k = 0; k < 4; k++) {
        const int i = (int)(curand_uniform(&st) * n_pts);
        uv[2 * k] = p2s[2 * i], uv[2 * k + 1] = p2s[2 * i + 1];
        X[k] = Vec3f{p3s[3 * i], p3s[3 * i + 1], p3s[3 * i + 2]};
    }
    Pose P[4];
    bool exists[4];
    float err[4];
    for (int q = 0; q < 4; q++) {
        err[q] = 0.f;
        exists[q] = solver == 0 ? twist_lane(q, uv, X, fx, fy, cx, cy, P[q], err[q]) : ap3p_lane(q, uv, X, fx, fy, cx, cy, P[q], err[q]);
    }
    const int best = pick_by_fourth_point(exists, err);
    const float nan = __int_as_float(0x7fffffff);
    if (best < 0) {
        for (int k = 0; k < 3; k++) rvecs[h * 3 + k] = nan, tvecs[h * 3 + k] = nan;
        return;
    }
    float R[3][3], rv[3];
    for (int k = 0; k < 9; k++) R[k / 3][k % 3] = P[best].R[k];
    vb::rot::rotation_to_rvec(R, rv);
    for (int k = 0; k < 3; k++) rvecs[h * 3 + k] = rv[k], tvecs[h * 3 + k] = P[best].t[k];
}

extern "C" unsigned probe_default_sites(int solver) { return solver == 0 ? kTwistSitesFused : kAp3pSitesFused; }

extern "C" int probe_solve(int solver, unsigned mask, const float* h_p2s, const float* h_p3s, int n_pts, int n_poses, float fx,
                           float fy, float cx, float cy, float* h_rvecs, float* h_tvecs) {
    float *p2 = nullptr, *p3 = nullptr, *rv = nullptr, *tv = nullptr;
    cudaMalloc((void**)&p2, (size_t)(n_pts + 1) * 2 * sizeof(float));
    cudaMalloc((void**)&p3, (size_t)(n_pts + 1) * 3 * sizeof(float));
    cudaMalloc((void**)&rv, (size_t)n_poses * 3 * sizeof(float));
    cudaMalloc((void**)&tv, (size_t)n_poses * 3 * sizeof(float));
    cudaMemset(p2, 0, (size_t)(n_pts + 1) * 2 * sizeof(float));
    cudaMemset(p3, 0, (size_t)(n_pts + 1) * 3 * sizeof(float));
    cudaMemcpy(p2, h_p2s, (size_t)n_pts * 2 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(p3, h_p3s, (size_t)n_pts * 3 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpyToSymbol(solver == 0 ? c_twist_sites : c_ap3p_sites, &mask, sizeof(mask));
    const unsigned other = solver == 0 ? kAp3pSitesFused : kTwistSitesFused;
    cudaMemcpyToSymbol(solver == 0 ? c_ap3p_sites : c_twist_sites, &other, sizeof(other));
    k_probe<<<(n_poses + 63) / 64, 64>>>(solver, p2, p3, n_pts, n_poses, fx, fy, cx, cy, rv, tv);
    cudaMemcpy(h_rvecs, rv, (size_t)n_poses * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    const cudaError_t e = cudaMemcpy(h_tvecs, tv, (size_t)n_poses * 3 * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(p2), cudaFree(p3), cudaFree(rv), cudaFree(tv);
    return (int)e;
}
