// Rotation matrix -> rotation vector on the device, through the McAdams et al. branch-free 3x3 SVD.
//
// Behavioural source: reference gpu-kernels/rodrigues.h:5-113 (orthonormalise R as U*V^T, then the
// Ceres-style matrix->angle-axis with atan2f) and gpu-kernels/svd3_cuda.h:36-1044 (A. McAdams, A. Selle,
// R. Tamstorf, J. Teran, E. Sifakis, "Computing the SVD of 3x3 matrices with minimal branching and
// elementary floating point operations", UW-Madison TR1690, 2011; 4 Jacobi sweeps, IEEE rsqrt + one Newton
// step).  The published algorithm is restated here with its three cyclic Jacobi conjugations, the three
// column sorts and the three Givens rotations expressed as calls of one helper each instead of unrolled
// text; every operation is an individually rounded FP32 op in the order of the reference.
#pragma once
#include "residual_model.cuh"  // f_mul / f_add / f_sub helpers

namespace vb {
namespace rot {

constexpr float kFourGammaSquared = 5.8284273147583007813f;  // (sqrt(8)+3)
constexpr float kSinPi8 = 0.3826834261417388916015625f;       // bits 0x3EC3EF15 (1053028117)
constexpr float kCosPi8 = 0.923879563808441162109375f;        // bits 0x3F6C835F (1064076127)
constexpr float kTiny = 1.e-20f;
constexpr float kSmall = 1.e-12f;

// rsqrt(x) with one Newton-Raphson step, in the operation order of the reference
__device__ __forceinline__ float rsqrt_nr(float x) {
    float r = __frsqrt_rn(x);
    const float h = f_mul(r, 0.5f);
    float t = f_mul(r, h);
    t = f_mul(r, t);
    t = f_mul(x, t);
    r = f_add(r, h);
    return f_sub(r, t);
}

// One approximate-Givens Jacobi conjugation of the symmetric matrix S on the (p,q) pair whose
// off-diagonal is s_pq; (s_rp, s_rq) is the remaining row, s_rr the remaining diagonal.  The quaternion
// (qp, qq, qr, qs) accumulates V.  Called with the three cyclic index assignments.
__device__ __forceinline__ void jacobi_conjugate(float& s_pp, float& s_pq, float& s_qq, float& s_rp, float& s_rq,
                                                 float& s_rr, float& q_p, float& q_q, float& q_r, float& q_s) {
    float sh = f_mul(s_pq, 0.5f);
    float t5 = f_sub(s_pp, s_qq);
    float t2 = f_mul(sh, sh);
    const bool big = (t2 >= kTiny);
    sh = big ? sh : 0.f;
    float ch = big ? t5 : 1.f;

    float t1 = f_mul(sh, sh);
    t2 = f_mul(ch, ch);
    float t3 = f_add(t1, t2);
    float t4 = __frsqrt_rn(t3);
    sh = f_mul(t4, sh);
    ch = f_mul(t4, ch);
    t1 = f_mul(kFourGammaSquared, t1);
    const bool clamp = (t2 <= t1);
    sh = clamp ? kSinPi8 : sh;
    ch = clamp ? kCosPi8 : ch;

    t1 = f_mul(sh, sh);
    t2 = f_mul(ch, ch);
    const float c = f_sub(t2, t1);
    float s = f_mul(ch, sh);
    s = f_add(s, s);

    // conjugation (un-normalised rotation: scale by sh^2+ch^2 first)
    t3 = f_add(t1, t2);
    s_rr = f_mul(s_rr, t3);
    s_rp = f_mul(s_rp, t3);
    s_rq = f_mul(s_rq, t3);
    s_rr = f_mul(s_rr, t3);

    t1 = f_mul(s, s_rp);
    t2 = f_mul(s, s_rq);
    s_rp = f_mul(c, s_rp);
    s_rq = f_mul(c, s_rq);
    s_rp = f_add(t2, s_rp);
    s_rq = f_sub(s_rq, t1);

    t2 = f_mul(s, s);
    t1 = f_mul(s_qq, t2);
    t3 = f_mul(s_pp, t2);
    t4 = f_mul(c, c);
    s_pp = f_mul(s_pp, t4);
    s_qq = f_mul(s_qq, t4);
    s_pp = f_add(s_pp, t1);
    s_qq = f_add(s_qq, t3);
    t4 = f_sub(t4, t2);
    t2 = f_add(s_pq, s_pq);
    s_pq = f_mul(s_pq, t4);
    t4 = f_mul(c, s);
    t2 = f_mul(t2, t4);
    t5 = f_mul(t5, t4);
    s_pp = f_add(s_pp, t2);
    s_pq = f_sub(s_pq, t5);
    s_qq = f_sub(s_qq, t2);

    // quaternion accumulation
    t1 = f_mul(sh, q_p);
    t2 = f_mul(sh, q_q);
    t3 = f_mul(sh, q_r);
    sh = f_mul(sh, q_s);
    q_s = f_mul(ch, q_s);
    q_p = f_mul(ch, q_p);
    q_q = f_mul(ch, q_q);
    q_r = f_mul(ch, q_r);
    q_r = f_add(q_r, sh);
    q_s = f_sub(q_s, t3);
    q_p = f_add(q_p, t2);
    q_q = f_sub(q_q, t1);
}

// swap columns i,j of B and V when norm_i < norm_j, then negate column `neg` (keeps det V = +1)
__device__ __forceinline__ void sort_columns(float B[3][3], float V[3][3], float n[3], int i, int j, int neg) {
    const bool sw = n[i] < n[j];
    if (sw) {
        for (int r = 0; r < 3; r++) {
            float t = B[r][i];
            B[r][i] = B[r][j];
            B[r][j] = t;
            t = V[r][i];
            V[r][i] = V[r][j];
            V[r][j] = t;
        }
        const float t = n[i];
        n[i] = n[j];
        n[j] = t;
    }
    const float sgn = f_add(1.f, sw ? -2.f : 0.f);
    for (int r = 0; r < 3; r++) {
        B[r][neg] = f_mul(B[r][neg], sgn);
        V[r][neg] = f_mul(V[r][neg], sgn);
    }
}

// Givens rotation zeroing B[q][col] against pivot B[p][col]; applied to rows p,q of B and columns p,q of U
__device__ __forceinline__ void qr_givens(float B[3][3], float U[3][3], int p, int q, int col) {
    const float apiv = B[p][col];
    const float abel = B[q][col];
    float sh = f_mul(abel, abel);
    sh = (sh >= kSmall) ? abel : 0.f;
    float ch = f_sub(0.f, apiv);
    ch = fmaxf(ch, apiv);
    ch = fmaxf(ch, kSmall);
    const bool nonneg = (apiv >= 0.f);

    float t1 = f_mul(ch, ch);
    float t2 = f_mul(sh, sh);
    t2 = f_add(t1, t2);
    t1 = rsqrt_nr(t2);
    t1 = f_mul(t1, t2);
    ch = f_add(ch, t1);
    if (!nonneg) {
        const float t = ch;
        ch = sh;
        sh = t;
    }
    t1 = f_mul(ch, ch);
    t2 = f_mul(sh, sh);
    t2 = f_add(t1, t2);
    t1 = rsqrt_nr(t2);
    ch = f_mul(ch, t1);
    sh = f_mul(sh, t1);
    float c = f_mul(ch, ch);
    float s = f_mul(sh, sh);
    c = f_sub(c, s);
    s = f_mul(sh, ch);
    s = f_add(s, s);

    for (int j = 0; j < 3; j++) {
        const float u1 = f_mul(s, B[p][j]);
        const float u2 = f_mul(s, B[q][j]);
        B[p][j] = f_mul(c, B[p][j]);
        B[q][j] = f_mul(c, B[q][j]);
        B[p][j] = f_add(B[p][j], u2);
        B[q][j] = f_sub(B[q][j], u1);
    }
    for (int i = 0; i < 3; i++) {
        const float u1 = f_mul(s, U[i][p]);
        const float u2 = f_mul(s, U[i][q]);
        U[i][p] = f_mul(c, U[i][p]);
        U[i][q] = f_mul(c, U[i][q]);
        U[i][p] = f_add(U[i][p], u2);
        U[i][q] = f_sub(U[i][q], u1);
    }
}

// A = U * diag(S) * V^T   (only U and V are needed by the caller)
__device__ __forceinline__ void svd3(const float A[3][3], float U[3][3], float V[3][3]) {
    // normal equations S = A^T A (lower triangle)
    float s11 = f_add(f_mul(A[2][0], A[2][0]), f_add(f_mul(A[1][0], A[1][0]), f_mul(A[0][0], A[0][0])));
    float s21 = f_add(f_mul(A[2][1], A[2][0]), f_add(f_mul(A[1][1], A[1][0]), f_mul(A[0][1], A[0][0])));
    float s31 = f_add(f_mul(A[2][2], A[2][0]), f_add(f_mul(A[1][2], A[1][0]), f_mul(A[0][2], A[0][0])));
    float s22 = f_add(f_mul(A[2][1], A[2][1]), f_add(f_mul(A[1][1], A[1][1]), f_mul(A[0][1], A[0][1])));
    float s32 = f_add(f_mul(A[2][2], A[2][1]), f_add(f_mul(A[1][2], A[1][1]), f_mul(A[0][2], A[0][1])));
    float s33 = f_add(f_mul(A[2][2], A[2][2]), f_add(f_mul(A[1][2], A[1][2]), f_mul(A[0][2], A[0][2])));

    float qs = 1.f, qx = 0.f, qy = 0.f, qz = 0.f;
    for (int sweep = 0; sweep < 4; sweep++) {
        jacobi_conjugate(s11, s21, s22, s31, s32, s33, qx, qy, qz, qs);
        jacobi_conjugate(s22, s32, s33, s21, s31, s11, qy, qz, qx, qs);
        jacobi_conjugate(s33, s31, s11, s32, s21, s22, qz, qx, qy, qs);
    }

    // normalise the quaternion
    float t2 = f_mul(qs, qs);
    t2 = f_add(f_mul(qx, qx), t2);
    t2 = f_add(f_mul(qy, qy), t2);
    t2 = f_add(f_mul(qz, qz), t2);
    const float rn = rsqrt_nr(t2);
    qs = f_mul(qs, rn);
    qx = f_mul(qx, rn);
    qy = f_mul(qy, rn);
    qz = f_mul(qz, rn);

    // quaternion -> V
    {
        const float xx = f_mul(qx, qx), yy = f_mul(qy, qy), zz = f_mul(qz, qz);
        float v11 = f_mul(qs, qs);
        float v22 = f_sub(v11, xx);
        float v33 = f_sub(v22, yy);
        v33 = f_add(v33, zz);
        v22 = f_add(v22, yy);
        v22 = f_sub(v22, zz);
        v11 = f_add(v11, xx);
        v11 = f_sub(v11, yy);
        v11 = f_sub(v11, zz);
        const float x2 = f_add(qx, qx), y2 = f_add(qy, qy), z2 = f_add(qz, qz);
        float v32 = f_mul(qs, x2);
        float v13 = f_mul(qs, y2);
        float v21 = f_mul(qs, z2);
        const float a = f_mul(qy, x2);
        const float b = f_mul(qz, y2);
        const float c = f_mul(qx, z2);
        V[0][1] = f_sub(a, v21);
        V[1][2] = f_sub(b, v32);
        V[2][0] = f_sub(c, v13);
        V[1][0] = f_add(a, v21);
        V[2][1] = f_add(b, v32);
        V[0][2] = f_add(c, v13);
        V[0][0] = v11, V[1][1] = v22, V[2][2] = v33;
    }

    // B = A * V
    float B[3][3];
    for (int i = 0; i < 3; i++) {
        const float a1 = A[i][0], a2 = A[i][1], a3 = A[i][2];
        float b2 = f_mul(V[0][1], a1);
        float b3 = f_mul(V[0][2], a1);
        float b1 = f_mul(V[0][0], a1);
        b1 = f_add(b1, f_mul(V[1][0], a2));
        b1 = f_add(b1, f_mul(V[2][0], a3));
        b2 = f_add(b2, f_mul(V[1][1], a2));
        b2 = f_add(b2, f_mul(V[2][1], a3));
        b3 = f_add(b3, f_mul(V[1][2], a2));
        b3 = f_add(b3, f_mul(V[2][2], a3));
        B[i][0] = b1, B[i][1] = b2, B[i][2] = b3;
    }

    // sort singular values (column norms), keeping V a rotation
    float n[3];
    for (int j = 0; j < 3; j++)
        n[j] = f_add(f_add(f_mul(B[0][j], B[0][j]), f_mul(B[1][j], B[1][j])), f_mul(B[2][j], B[2][j]));
    sort_columns(B, V, n, 0, 1, 1);
    sort_columns(B, V, n, 0, 2, 0);
    sort_columns(B, V, n, 1, 2, 2);

    // QR of B by Givens rotations -> U
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) U[i][j] = (i == j) ? 1.f : 0.f;
    qr_givens(B, U, 0, 1, 0);
    qr_givens(B, U, 0, 2, 0);
    qr_givens(B, U, 1, 2, 1);
}

// rotation matrix -> angle-axis (Ceres RotationMatrixToAngleAxis with atan2f), reference rodrigues.h:5-79
__device__ __forceinline__ void matrix_to_angle_axis(const float R[3][3], float aa[3]) {
    aa[0] = R[2][1] - R[1][2];
    aa[1] = R[0][2] - R[2][0];
    aa[2] = R[1][0] - R[0][1];
    const float costheta = fminf(fmaxf((R[0][0] + R[1][1] + R[2][2] - 1.f) * 0.5f, -1.f), 1.f);
    const float sintheta = fminf(sqrtf(aa[0] * aa[0] + aa[1] * aa[1] + aa[2] * aa[2]) * 0.5f, 1.f);
    const float theta = atan2f(sintheta, costheta);
    if ((sintheta > FLT_EPSILON) || (sintheta < -FLT_EPSILON)) {
        const float r = theta / (2.f * sintheta);
        aa[0] *= r;
        aa[1] *= r;
        aa[2] *= r;
        return;
    }
    if (costheta > 0) {
        aa[0] *= 0.5f;
        aa[1] *= 0.5f;
        aa[2] *= 0.5f;
        return;
    }
    const float inv_one_minus_costheta = 1.f / (1.f - costheta);
    for (int i = 0; i < 3; ++i) {
        aa[i] = theta * sqrtf((R[i][i] - costheta) * inv_one_minus_costheta);
        if (((sintheta < 0) && (aa[i] > 0)) || ((sintheta > 0) && (aa[i] < 0))) aa[i] = -aa[i];
    }
}

// R (approximately a rotation) -> rvec: project onto SO(3) as U*V^T, then angle-axis (rodrigues.h:82-113)
__device__ __forceinline__ void rotation_to_rvec(const float Rin[3][3], float rvec[3]) {
    float U[3][3], V[3][3], R[3][3];
    svd3(Rin, U, V);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R[i][j] = U[i][0] * V[j][0] + U[i][1] * V[j][1] + U[i][2] * V[j][2];
    matrix_to_angle_axis(R, rvec);
}

}  // namespace rot
}  // namespace vb
