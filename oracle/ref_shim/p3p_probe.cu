// TEST INFRASTRUCTURE (oracle). Diagnostic probe: evaluates the reference's lambdatwist / rodrigues device
// functions (headers taken from /root/reference by -I at build time) and ours on the same inputs, stage by
// stage, and counts bit mismatches per stage.  Used to localise rounding differences; not a parity test.
#include "rodrigues.h"                       // reference gpu-kernels/rodrigues.h (+ svd3_cuda.h)
#include "../lambdatwist/lambdatwist_p4p.h"  // reference lambdatwist
#undef gone
#include "../../voldor_b200/csrc/p3p_lambdatwist.cuh"
#include "../../voldor_b200/csrc/rotation.cuh"
#include <cstdio>

namespace {
__device__ inline bool beq(float a, float b) { return __float_as_uint(a) == __float_as_uint(b); }

// stage ids: 0 cubic, 1 eig, 2 refine, 3 p3p valid count, 4 p3p R/T, 5 p4p R/t, 6 rodrigues
__global__ void k_probe(const float* p2s, const float* p3s, const float4* u4, int n_pts, int n_poses, float fx, float fy,
                        float cx, float cy, unsigned long long* mism, float* dump) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_poses) return;
    const float4 u = u4[idx];
    const int i1 = (int)(u.x * n_pts), i2 = (int)(u.y * n_pts), i3 = (int)(u.z * n_pts), i4 = (int)(u.w * n_pts);
    const float *y1 = &p2s[i1 * 2], *y2 = &p2s[i2 * 2], *y3 = &p2s[i3 * 2], *y4 = &p2s[i4 * 2];
    const float *x1 = &p3s[i1 * 3], *x2 = &p3s[i2 * 3], *x3 = &p3s[i3 * 3], *x4 = &p3s[i4 * 3];

    // ---- stage 4/3: full p3p
    cvl::Vector3<float> vy1((y1[0] - cx) / fx, (y1[1] - cy) / fy, 1.0f), vy2((y2[0] - cx) / fx, (y2[1] - cy) / fy, 1.0f),
        vy3((y3[0] - cx) / fx, (y3[1] - cy) / fy, 1.0f);
    cvl::Vector3<float> vx1(x1[0], x1[1], x1[2]), vx2(x2[0], x2[1], x2[2]), vx3(x3[0], x3[1], x3[2]);
    cvl::Vector<cvl::Matrix<float, 3, 3>, 4> rRs;
    cvl::Vector<cvl::Vector3<float>, 4> rTs;
    const int nr = cvl::p3p_lambdatwist<float, 5>(vy1, vy2, vy3, vx1, vx2, vx3, rRs, rTs);

    using namespace vb::p3p;
    Mat3 mRs[4];
    Vec3 mTs[4];
    const int nm = p3p_solve<5>(make_vec3((y1[0] - cx) / fx, (y1[1] - cy) / fy, 1.0f),
                                make_vec3((y2[0] - cx) / fx, (y2[1] - cy) / fy, 1.0f),
                                make_vec3((y3[0] - cx) / fx, (y3[1] - cy) / fy, 1.0f), make_vec3(x1[0], x1[1], x1[2]),
                                make_vec3(x2[0], x2[1], x2[2]), make_vec3(x3[0], x3[1], x3[2]), mRs, mTs);
    if (nr != nm) atomicAdd(&mism[3], 1ULL);
    bool rt_bad = false;
    for (int i = 0; i < min(nr, nm); i++) {
        for (int k = 0; k < 9; k++)
            if (!beq(rRs(i)(k), mRs[i].m[k])) rt_bad = true;
        for (int k = 0; k < 3; k++)
            if (!beq(rTs(i)(k), mTs[i][k])) rt_bad = true;
    }
    if (rt_bad) atomicAdd(&mism[4], 1ULL);

    // ---- stage 0..2 on derived inputs (cheap pseudo-inputs from the data keep magnitudes realistic)
    {
        const float b = x1[0] - x2[1] * 0.37f, c = x3[2] * 0.21f - x1[1], d = (y1[0] - y2[0]) * 0.003f;
        const float gr = cvl::cubick<float>(b, c, d), gm = cubic_root(b, c, d);
        if (!beq(gr, gm)) atomicAdd(&mism[0], 1ULL);
    }
    {
        // symmetric matrix with (near) zero eigenvalue: A = a a^T + b b^T
        const float a0 = x1[0], a1 = x1[1], a2 = x1[2] * 0.1f, b0 = x2[0] * 0.3f, b1 = -x2[1], b2 = x2[2] * 0.05f;
        cvl::Matrix<float, 3, 3> A(a0 * a0 + b0 * b0, a0 * a1 + b0 * b1, a0 * a2 + b0 * b2, a0 * a1 + b0 * b1,
                                   a1 * a1 + b1 * b1, a1 * a2 + b1 * b2, a0 * a2 + b0 * b2, a1 * a2 + b1 * b2,
                                   a2 * a2 + b2 * b2);
        cvl::Matrix<float, 3, 3> V;
        cvl::Vector3<float> L;
        cvl::eigwithknown0(A, V, L);
        Mat3 Am, Vm;
        Vec3 Lm;
        for (int k = 0; k < 9; k++) Am.m[k] = A(k);
        eig_known0(Am, Vm, Lm);
        bool bad = false;
        for (int k = 0; k < 9; k++)
            if (!beq(V(k), Vm.m[k])) bad = true;
        for (int k = 0; k < 3; k++)
            if (!beq(L(k), Lm[k])) bad = true;
        if (bad) atomicAdd(&mism[1], 1ULL);
    }
    {
        const float a12 = 1.0f + fabsf(x1[0]), a13 = 0.8f + fabsf(x2[1]), a23 = 1.2f + fabsf(x3[0]);
        const float b12 = -1.9f, b13 = -1.8f + 0.01f * y1[0] / fx, b23 = -1.95f;
        cvl::Vector3<float> Lr(x1[2], x2[2], x3[2]);
        Vec3 Lm = make_vec3(x1[2], x2[2], x3[2]);
        cvl::gauss_newton_refineL<float, 5>(Lr, a12, a13, a23, b12, b13, b23);
        refine_lambda<5>(Lm, a12, a13, a23, b12, b13, b23);
        bool bad = false;
        for (int k = 0; k < 3; k++)
            if (!beq(Lr(k), Lm[k])) bad = true;
        if (bad) atomicAdd(&mism[2], 1ULL);
    }
    // ---- stage 5: p4p
    float Rr[3][3], tr[3], Rm[3][3], tm[3];
    const bool okr = lambdatwist_p4p<float, float, 5>((float*)y1, (float*)y2, (float*)y3, (float*)y4, (float*)x1, (float*)x2,
                                                      (float*)x3, (float*)x4, fx, fy, cx, cy, Rr, tr);
    const bool okm = p4p_solve(y1, y2, y3, y4, x1, x2, x3, x4, fx, fy, cx, cy, Rm, tm);
    bool bad5 = okr != okm;
    if (okr && okm) {
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++)
                if (!beq(Rr[r][c], Rm[r][c])) bad5 = true;
        for (int k = 0; k < 3; k++)
            if (!beq(tr[k], tm[k])) bad5 = true;
    }
    if (bad5) atomicAdd(&mism[5], 1ULL);
    // ---- stage 6: rodrigues on the reference's R
    if (okr) {
        float R1[3][3], rv1[3], rv2[3];
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) R1[r][c] = Rr[r][c];
        vb::rot::rotation_to_rvec(Rr, rv2);
        rodrigues(R1, rv1);
        if (!beq(rv1[0], rv2[0]) || !beq(rv1[1], rv2[1]) || !beq(rv1[2], rv2[2])) atomicAdd(&mism[6], 1ULL);
        if (idx < 64) {
            for (int k = 0; k < 3; k++) dump[idx * 6 + k] = rv1[k], dump[idx * 6 + 3 + k] = rv2[k];
        }
    }
}

// "clean" contexts: one solver per kernel, like the production kernels
__global__ void k_clean_ref(const float* p2s, const float* p3s, const float4* u4, int n_pts, int n_poses, float fx,
                            float fy, float cx, float cy, float* out /*[n][6] rvec,t*/) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_poses) return;
    const float4 u = u4[idx];
    const int i1 = (int)(u.x * n_pts), i2 = (int)(u.y * n_pts), i3 = (int)(u.z * n_pts), i4 = (int)(u.w * n_pts);
    float R[3][3], t[3];
    const bool ok = lambdatwist_p4p<float, float, 5>((float*)&p2s[i1 * 2], (float*)&p2s[i2 * 2], (float*)&p2s[i3 * 2],
                                                     (float*)&p2s[i4 * 2], (float*)&p3s[i1 * 3], (float*)&p3s[i2 * 3],
                                                     (float*)&p3s[i3 * 3], (float*)&p3s[i4 * 3], fx, fy, cx, cy, R, t);
    float* o = out + idx * 6;
    if (!ok) {
        for (int k = 0; k < 6; k++) o[k] = CUDART_NAN_F;
        return;
    }
    o[3] = t[0], o[4] = t[1], o[5] = t[2];
    rodrigues(R, o);
}
__global__ void k_clean_mine(const float* p2s, const float* p3s, const float4* u4, int n_pts, int n_poses, float fx,
                             float fy, float cx, float cy, float* out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_poses) return;
    const float4 u = u4[idx];
    const int i1 = (int)(u.x * n_pts), i2 = (int)(u.y * n_pts), i3 = (int)(u.z * n_pts), i4 = (int)(u.w * n_pts);
    float R[3][3], t[3];
    const bool ok = vb::p3p::p4p_solve(&p2s[i1 * 2], &p2s[i2 * 2], &p2s[i3 * 2], &p2s[i4 * 2], &p3s[i1 * 3], &p3s[i2 * 3],
                                       &p3s[i3 * 3], &p3s[i4 * 3], fx, fy, cx, cy, R, t);
    float* o = out + idx * 6;
    if (!ok) {
        for (int k = 0; k < 6; k++) o[k] = CUDART_NAN_F;
        return;
    }
    o[3] = t[0], o[4] = t[1], o[5] = t[2];
    vb::rot::rotation_to_rvec(R, o);
}

__global__ void k_draws(float4* u4, int n) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    curandStateXORWOW_t st;
    curand_init(233ULL, (unsigned long long)idx, 0, &st);
    u4[idx] = make_float4(curand_uniform(&st), curand_uniform(&st), curand_uniform(&st), curand_uniform(&st));
}
}  // namespace

extern "C" int p3p_probe(const float* h_p3s, const float* h_p2s, const float* h_K, int n_pts, int n_poses,
                         unsigned long long* h_mism /*[8]*/, float* h_dump /*[64*6]*/, float* h_clean_ref,
                         float* h_clean_mine) {
    float *p2, *p3, *dump;
    float4* u4;
    unsigned long long* mism;
    cudaMalloc(&p2, (n_pts + 1) * 2 * sizeof(float)), cudaMalloc(&p3, (n_pts + 1) * 3 * sizeof(float));
    cudaMalloc(&u4, n_poses * sizeof(float4)), cudaMalloc(&mism, 8 * sizeof(unsigned long long));
    cudaMalloc(&dump, 64 * 6 * sizeof(float));
    cudaMemset(mism, 0, 64), cudaMemset(dump, 0, 64 * 6 * sizeof(float));
    cudaMemset(p2, 0, (n_pts + 1) * 2 * sizeof(float)), cudaMemset(p3, 0, (n_pts + 1) * 3 * sizeof(float));
    cudaMemcpy(p2, h_p2s, n_pts * 2 * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemcpy(p3, h_p3s, n_pts * 3 * sizeof(float), cudaMemcpyHostToDevice);
    k_draws<<<(n_poses + 127) / 128, 128>>>(u4, n_poses);
    k_probe<<<(n_poses + 31) / 32, 32>>>(p2, p3, u4, n_pts, n_poses, h_K[0], h_K[4], h_K[2], h_K[5], mism, dump);
    {
        float* o;
        cudaMalloc(&o, n_poses * 6 * sizeof(float));
        k_clean_ref<<<(n_poses + 31) / 32, 32>>>(p2, p3, u4, n_pts, n_poses, h_K[0], h_K[4], h_K[2], h_K[5], o);
        cudaMemcpy(h_clean_ref, o, n_poses * 6 * sizeof(float), cudaMemcpyDeviceToHost);
        k_clean_mine<<<(n_poses + 31) / 32, 32>>>(p2, p3, u4, n_pts, n_poses, h_K[0], h_K[4], h_K[2], h_K[5], o);
        cudaMemcpy(h_clean_mine, o, n_poses * 6 * sizeof(float), cudaMemcpyDeviceToHost);
        cudaFree(o);
    }
    cudaMemcpy(h_mism, mism, 64, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_dump, dump, 64 * 6 * sizeof(float), cudaMemcpyDeviceToHost);
    const int rc = (int)cudaGetLastError();
    cudaFree(p2), cudaFree(p3), cudaFree(u4), cudaFree(mism), cudaFree(dump);
    return rc;
}
