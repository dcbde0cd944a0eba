// AP3P minimal solver (T. Ke, S. Roumeliotis, "An Efficient Algebraic Solution to the Perspective-Three-
// Point Problem", CVPR 2017) + 4th-point disambiguation, FP32.
//
// Behavioural source: reference gpu-kernels/solve_batch_ap3p.cu:9-26 (complex helpers), :28-82 (Ferrari
// quartic), :85-99 (2 Newton polish passes), :152-292 (pose recovery), :294-328 (bearing normalisation),
// :331-378 (hypothesis kernel body: solution choice by reprojection of the 4th point, no cheirality test).
// Selected by the reference with `--lambdatwist 0` (voldor/geometry.cpp:149-154).
#pragma once
#include <cuComplex.h>
#include <cuda_runtime.h>
#include <math.h>

namespace vb {
namespace ap3p {

// principal square root with non-positive imaginary part (solve_batch_ap3p.cu:9-15)
__device__ inline cuFloatComplex csqrt_neg_imag(cuFloatComplex x) {
    cuFloatComplex out;
    out.x = sqrtf(cuCabsf(x) * (x.x / cuCabsf(x) + 1.0f) / 2.0f);
    out.y = sqrtf(cuCabsf(x) * (1.0f - x.x / cuCabsf(x)) / 2.0f);
    out.y = -fabsf(out.y);
    return out;
}

__device__ inline cuFloatComplex cpow_real(const cuFloatComplex& z, float p) {
    const float theta = atan2f(z.y, z.x);
    return make_cuFloatComplex((powf(cuCabsf(z), p) * cosf(p * theta)), (powf(cuCabsf(z), p) * sinf(p * theta)));
}

__device__ inline cuFloatComplex cneg(cuFloatComplex x) { return make_cuFloatComplex(-x.x, -x.y); }

// Ferrari's closed form for a4 x^4 + a3 x^3 + a2 x^2 + a1 x + a0 (real parts of the four roots)
__device__ inline void solve_quartic(const float* factors, float* realRoots) {
    const float a4 = factors[0], a3 = factors[1], a2 = factors[2], a1 = factors[3], a0 = factors[4];

    const float a4_2 = a4 * a4;
    const float a3_2 = a3 * a3;
    const float a4_3 = a4_2 * a4;
    const float a2a4 = a2 * a4;

    const float p4 = (8 * a2a4 - 3 * a3_2) / (8 * a4_2);
    const float q4 = (a3_2 * a3 - 4 * a2a4 * a3 + 8 * a1 * a4_2) / (8 * a4_3);
    const float r4 =
        (256 * a0 * a4_3 - 3 * (a3_2 * a3_2) - 64 * a1 * a3 * a4_2 + 16 * a2a4 * a3_2) / (256 * (a4_3 * a4));

    const float p3 = ((p4 * p4) / 12 + r4) / 3;
    const float q3 = (72 * r4 * p4 - 2 * p4 * p4 * p4 - 27 * q4 * q4) / 432;

    float t;
    cuFloatComplex w = make_cuFloatComplex(q3 * q3 - p3 * p3 * p3, 0);
    w = csqrt_neg_imag(w);
    if (q3 >= 0) {
        w.x = -w.x - q3;
        w.y = -w.y;
    } else {
        w = csqrt_neg_imag(w);  // the reference takes the root twice on this branch
        w.x = w.x - q3;
    }
    if (w.y == 0.0f) {
        w.x = cbrtf(w.x);
        t = 2.0f * (w.x + p3 / w.x);
    } else {
        w = cpow_real(w, (1.0f / 3.0f));
        t = 4.0f * w.x;
    }

    const cuFloatComplex sqrt_2m = csqrt_neg_imag(make_cuFloatComplex(-2 * p4 / 3 + t, 0));
    const float B_4A = -a3 / (4 * a4);
    const cuFloatComplex complex1 = make_cuFloatComplex(4 * p4 / 3 + t, 0);
    const cuFloatComplex complex2 = cuCdivf(make_cuFloatComplex(2 * q4, 0), sqrt_2m);

    const float sqrt_2m_rh = sqrt_2m.x * 0.5f;
    const float sqrt1 = csqrt_neg_imag(cneg(cuCaddf(complex1, complex2))).x * 0.5f;
    realRoots[0] = B_4A + sqrt_2m_rh + sqrt1;
    realRoots[1] = B_4A + sqrt_2m_rh - sqrt1;
    const float sqrt2 = csqrt_neg_imag(cneg(cuCsubf(complex1, complex2))).x * 0.5f;
    realRoots[2] = B_4A - sqrt_2m_rh + sqrt2;
    realRoots[3] = B_4A - sqrt_2m_rh - sqrt2;
}

__device__ inline void polish_quartic_roots(const float* coeffs, float* roots) {
    for (int i = 0; i < 2; ++i) {
        for (int j = 0; j < 4; ++j) {
            const float error =
                (((coeffs[0] * roots[j] + coeffs[1]) * roots[j] + coeffs[2]) * roots[j] + coeffs[3]) * roots[j] +
                coeffs[4];
            const float derivative =
                ((4 * coeffs[0] * roots[j] + 3 * coeffs[1]) * roots[j] + 2 * coeffs[2]) * roots[j] + coeffs[3];
            roots[j] -= error / derivative;
        }
    }
}

__device__ inline void v_cross(const float* a, const float* b, float* r) {
    r[0] = a[1] * b[2] - a[2] * b[1];
    r[1] = -(a[0] * b[2] - a[2] * b[0]);
    r[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline float v_dot(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ inline float v_norm(const float* a) { return sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }
__device__ inline void v_scale(float s, const float* a, float* r) {
    r[0] = a[0] * s, r[1] = a[1] * s, r[2] = a[2] * s;
}
__device__ inline void v_sub(const float* a, const float* b, float* r) {
    r[0] = a[0] - b[0], r[1] = a[1] - b[1], r[2] = a[2] - b[2];
}
__device__ inline void v_div(const float* a, float d, float* r) { r[0] = a[0] / d, r[1] = a[1] / d, r[2] = a[2] / d; }
__device__ inline void m_mult(const float a[3][3], const float b[3][3], float r[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}
// same product with the contraction the reference build uses for R = (Ck1nl*C13)*Cb1k3tzT: first term a rounded
// multiply, the other two fused (verified in the reference's PTX; the compiler's choice is context dependent)
__device__ inline void m_mult_first_plain(const float a[3][3], const float b[3][3], float r[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            r[i][j] = __fmaf_rn(a[i][2], b[2][j], __fmaf_rn(a[i][1], b[1][j], __fmul_rn(a[i][0], b[0][j])));
}

// b*: unit bearing vectors, w*: world points.  Up to 4 (R, t) with X_cam = R X_world + t.
__device__ inline int compute_poses(const float b1[3], const float b2[3], const float b3[3], const float w1[3],
                                    const float w2[3], const float w3[3], float solR[4][3][3], float solT[4][3]) {
    float u0[3];
    v_sub(w1, w2, u0);
    const float nu0 = v_norm(u0);
    float k1[3];
    v_div(u0, nu0, k1);

    float k3[3];
    v_cross(b1, b2, k3);
    const float nk3 = v_norm(k3);
    v_div(k3, nk3, k3);
    float tz[3];
    v_cross(b1, k3, tz);

    float v1[3], v2[3];
    v_cross(b1, b3, v1);
    v_cross(b2, b3, v2);
    float u1[3];
    v_sub(w1, w3, u1);

    const float u1k1 = v_dot(u1, k1);
    const float k3b3 = v_dot(k3, b3);
    float f11 = k3b3;
    float f13 = v_dot(k3, v1);
    const float f15 = -u1k1 * f11;
    float nl[3];
    v_cross(u1, k1, nl);
    const float delta = v_norm(nl);
    v_div(nl, delta, nl);
    f11 *= delta;
    f13 *= delta;

    const float u2k1 = u1k1 - nu0;
    float f21 = v_dot(tz, v2);
    float f22 = nk3 * k3b3;
    float f23 = v_dot(k3, v2);
    const float f24 = u2k1 * f22;
    const float f25 = -u2k1 * f21;
    f21 *= delta;
    f22 *= delta;
    f23 *= delta;
    const float g1 = f13 * f22;
    const float g2 = f13 * f25 - f15 * f23;
    const float g3 = f11 * f23 - f13 * f21;
    const float g4 = -f13 * f24;
    const float g5 = f11 * f22;
    const float g6 = f11 * f25 - f15 * f21;
    const float g7 = -f15 * f24;
    const float coeffs[5] = {g5 * g5 + g1 * g1 + g3 * g3, 2 * (g5 * g6 + g1 * g2 + g3 * g4),
                             g6 * g6 + 2 * g5 * g7 + g2 * g2 + g4 * g4 - g1 * g1 - g3 * g3,
                             2 * (g6 * g7 - g1 * g2 - g3 * g4), g7 * g7 - g2 * g2 - g4 * g4};
    float s[4];
    solve_quartic(coeffs, s);
    polish_quartic_roots(coeffs, s);

    float temp[3];
    v_cross(k1, nl, temp);
    const float Ck1nl[3][3] = {{k1[0], nl[0], temp[0]}, {k1[1], nl[1], temp[1]}, {k1[2], nl[2], temp[2]}};
    const float Cb1k3tzT[3][3] = {{b1[0], b1[1], b1[2]}, {k3[0], k3[1], k3[2]}, {tz[0], tz[1], tz[2]}};
    float b3p[3];
    v_scale((delta / k3b3), b3, b3p);

    int nb = 0;
    for (int i = 0; i < 4; ++i) {
        const float ctheta1p = s[i];
        if (fabsf(ctheta1p) > 1) continue;
        float stheta1p = sqrtf(1 - ctheta1p * ctheta1p);
        stheta1p = (k3b3 > 0) ? stheta1p : -stheta1p;
        float ctheta3 = g1 * ctheta1p + g2;
        float stheta3 = g3 * ctheta1p + g4;
        const float ntheta3 = stheta1p / ((g5 * ctheta1p + g6) * ctheta1p + g7);
        ctheta3 *= ntheta3;
        stheta3 *= ntheta3;

        const float C13[3][3] = {{ctheta3, 0, -stheta3},
                                 {stheta1p * stheta3, ctheta1p, stheta1p * ctheta3},
                                 {ctheta1p * stheta3, -stheta1p, ctheta1p * ctheta3}};
        float tm[3][3], R[3][3];
        m_mult(Ck1nl, C13, tm);
        m_mult_first_plain(tm, Cb1k3tzT, R);

        const float rp3[3] = {w3[0] * R[0][0] + w3[1] * R[1][0] + w3[2] * R[2][0],
                              w3[0] * R[0][1] + w3[1] * R[1][1] + w3[2] * R[2][1],
                              w3[0] * R[0][2] + w3[1] * R[1][2] + w3[2] * R[2][2]};
        float pxs[3];
        v_scale(stheta1p, b3p, pxs);
        v_sub(pxs, rp3, solT[nb]);
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) solR[nb][r][c] = R[c][r];  // transpose (world -> camera)
        nb++;
    }
    return nb;
}

__device__ inline void bearing(float mu, float mv, float fx, float fy, float cx, float cy, float out[3]) {
    mu = (mu - cx) / fx;
    mv = (mv - cy) / fy;
    const float norm = sqrtf(mu * mu + mv * mv + 1);
    const float mk = 1.f / norm;
    out[0] = mu * mk, out[1] = mv * mk, out[2] = mk;
}

__device__ inline bool p4p_solve(const float* y1, const float* y2, const float* y3, const float* y4, const float* x1,
                                 const float* x2, const float* x3, const float* x4, float fx, float fy, float cx,
                                 float cy, float R[3][3], float t[3]) {
    float b1[3], b2[3], b3[3];
    bearing(y1[0], y1[1], fx, fy, cx, cy, b1);
    bearing(y2[0], y2[1], fx, fy, cx, cy, b2);
    bearing(y3[0], y3[1], fx, fy, cx, cy, b3);
    float Rs[4][3][3], ts[4][3];
    const int n = compute_poses(b1, b2, b3, x1, x2, x3, Rs, ts);
    if (n == 0) return false;
    int ns = 0;
    float min_reproj = 0;
    for (int i = 0; i < n; i++) {
        const float X3p = Rs[i][0][0] * x4[0] + Rs[i][0][1] * x4[1] + Rs[i][0][2] * x4[2] + ts[i][0];
        const float Y3p = Rs[i][1][0] * x4[0] + Rs[i][1][1] * x4[1] + Rs[i][1][2] * x4[2] + ts[i][1];
        const float Z3p = Rs[i][2][0] * x4[0] + Rs[i][2][1] * x4[1] + Rs[i][2][2] * x4[2] + ts[i][2];
        const float mu3p = cx + fx * X3p / Z3p;
        const float mv3p = cy + fy * Y3p / Z3p;
        const float reproj = (mu3p - y4[0]) * (mu3p - y4[0]) + (mv3p - y4[1]) * (mv3p - y4[1]);
        if (i == 0 || min_reproj > reproj) {
            ns = i;
            min_reproj = reproj;
        }
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r][c] = Rs[ns][r][c];
    t[0] = ts[ns][0], t[1] = ts[ns][1], t[2] = ts[ns][2];
    return true;
}

}  // namespace ap3p
}  // namespace vb
