// Small host-side math used by the window orchestration: rotation-vector conversions and norms that the
// reference takes from OpenCV (cv::Rodrigues, cv::norm), restated without OpenCV.
//
// Behavioural source: the reference calls cv::Rodrigues on float data at voldor/voldor.cpp:64,
// voldor/utils.h:52-56 and voldor/geometry.cpp:184,258; OpenCV evaluates both directions in double and
// narrows the result to float.  rvec->R is the Rodrigues formula; R->rvec first projects R onto SO(3)
// (U*V^T of its SVD) and then uses the (R - R^T)/2 axis with acos of the trace, like OpenCV's
// cvRodrigues2.  These are third-party arithmetic for the reference too: results agree with
// OpenCV to double rounding before the final float narrowing, they are not claimed bit-identical to it.
#pragma once
#include <cmath>
#include <cstring>

#if defined(__CUDACC__)
#define VB_HD __host__ __device__
#else
#define VB_HD
#endif

namespace vb {
namespace hm {

inline void mat3_mul(const double* A, const double* B, double* C) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

// ---------------------------------------------------------------------------------------------------------
// Deterministic double arithmetic shared by host and device code.  rvec -> R runs on the host (window set-up,
// oracle orchestration) AND inside the mean-shift kernel (the pose of camera i feeds camera i+1 of the same EM
// iteration without a host round trip), so it must give the same bits on both sides: every operation is an
// individually rounded IEEE add/mul/div/sqrt (explicit *_rn intrinsics on the device, no FMA contraction on the
// host: the translation units are built with -ffp-contract=off), and sin/cos are evaluated here instead of by
// libm / libdevice.
// ---------------------------------------------------------------------------------------------------------
namespace det {
VB_HD inline double mul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
VB_HD inline double add(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
VB_HD inline double sub(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
VB_HD inline double div(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    return a / b;
#endif
}
VB_HD inline double root(double a) {
#if defined(__CUDA_ARCH__)
    return __dsqrt_rn(a);
#else
    return std::sqrt(a);
#endif
}

// sin and cos of x >= 0: Cody-Waite reduction by pi/2 (two-part constant), then the classic minimax kernels on
// [-pi/4, pi/4] (coefficients of the freely distributable fdlibm k_sin.c / k_cos.c), Horner form.
VB_HD inline void sincos(double x, double* s_out, double* c_out) {
    const double n = floor(add(mul(x, 6.36619772367581382433e-01), 0.5));  // nearest multiple of pi/2
    const double r = sub(sub(x, mul(n, 1.57079632673412561417e+00)), mul(n, 6.07710050650619224932e-11));
    const double z = mul(r, r);
    // sin(r)
    double p = add(-2.50507602534068634195e-08, mul(z, 1.58969099521155010221e-10));
    p = add(2.75573137070700676789e-06, mul(z, p));
    p = add(-1.98412698298579493134e-04, mul(z, p));
    p = add(8.33333333332248946124e-03, mul(z, p));
    p = add(-1.66666666666666324348e-01, mul(z, p));
    const double sr = add(r, mul(mul(z, r), p));
    // cos(r)
    double q = add(2.08757232129817482790e-09, mul(z, -1.13596475577881948265e-11));
    q = add(-2.75573143513906633035e-07, mul(z, q));
    q = add(2.48015872894767294178e-05, mul(z, q));
    q = add(-1.38888888888741095749e-03, mul(z, q));
    q = add(4.16666666666666019037e-02, mul(z, q));
    const double cr = sub(1.0, sub(mul(0.5, z), mul(mul(z, z), q)));
    const int quad = (int)(n - 4.0 * floor(n * 0.25));  // exact: n is a small non-negative integer
    double s = sr, c = cr;
    if (quad == 1) s = cr, c = -sr;
    if (quad == 2) s = -sr, c = -cr;
    if (quad == 3) s = -cr, c = sr;
    *s_out = s, *c_out = c;
}
}  // namespace det

// rvec (float[3]) -> R (float[9], row-major); host and device, bit-identical on both
VB_HD inline void rvec_to_matrix(const float* rvec, float* R9) {
    using namespace det;
    const double rx = rvec[0], ry = rvec[1], rz = rvec[2];
    const double theta = root(add(add(mul(rx, rx), mul(ry, ry)), mul(rz, rz)));
    double R[9];
    if (theta < 2.220446049250313e-16) {
        R[0] = 1, R[1] = 0, R[2] = 0, R[3] = 0, R[4] = 1, R[5] = 0, R[6] = 0, R[7] = 0, R[8] = 1;
    } else {
        double s, c;
        sincos(theta, &s, &c);
        const double c1 = sub(1., c);
        const double itheta = div(1., theta);
        const double x = mul(rx, itheta), y = mul(ry, itheta), z = mul(rz, itheta);
        const double c1x = mul(c1, x), c1y = mul(c1, y), c1z = mul(c1, z);
        const double sx = mul(s, x), sy = mul(s, y), sz = mul(s, z);
        R[0] = add(c, mul(c1x, x)), R[1] = sub(mul(c1x, y), sz), R[2] = add(mul(c1x, z), sy);
        R[3] = add(mul(c1x, y), sz), R[4] = add(c, mul(c1y, y)), R[5] = sub(mul(c1y, z), sx);
        R[6] = sub(mul(c1x, z), sy), R[7] = add(mul(c1y, z), sx), R[8] = add(c, mul(c1z, z));
    }
    for (int i = 0; i < 9; i++) R9[i] = (float)R[i];
}

// nearest rotation U*V^T of a 3x3 (double) by one-sided Jacobi on A^T A
inline void orthonormalize(const double* A, double* Q) {
    // V from the symmetric eigenproblem of S = A^T A by cyclic Jacobi
    double S[9], V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) S[i * 3 + j] = A[i] * A[j] + A[3 + i] * A[3 + j] + A[6 + i] * A[6 + j];
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = std::fabs(S[1]) + std::fabs(S[2]) + std::fabs(S[5]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                const double apq = S[p * 3 + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double tau = (S[q * 3 + q] - S[p * 3 + p]) / (2 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
                const double c = 1 / std::sqrt(1 + t * t), s = t * c;
                for (int k = 0; k < 3; k++) {  // S <- S J
                    const double skp = S[k * 3 + p], skq = S[k * 3 + q];
                    S[k * 3 + p] = c * skp - s * skq;
                    S[k * 3 + q] = s * skp + c * skq;
                }
                for (int k = 0; k < 3; k++) {  // S <- J^T S
                    const double spk = S[p * 3 + k], sqk = S[q * 3 + k];
                    S[p * 3 + k] = c * spk - s * sqk;
                    S[q * 3 + k] = s * spk + c * sqk;
                }
                for (int k = 0; k < 3; k++) {
                    const double vkp = V[k * 3 + p], vkq = V[k * 3 + q];
                    V[k * 3 + p] = c * vkp - s * vkq;
                    V[k * 3 + q] = s * vkp + c * vkq;
                }
            }
    }
    // Q = A V diag(1/sigma) V^T
    double AV[9];
    mat3_mul(A, V, AV);
    double W[9];
    for (int j = 0; j < 3; j++) {
        const double n = std::sqrt(AV[j] * AV[j] + AV[3 + j] * AV[3 + j] + AV[6 + j] * AV[6 + j]);
        const double inv = n > 0 ? 1.0 / n : 0.0;
        for (int i = 0; i < 3; i++) W[i * 3 + j] = AV[i * 3 + j] * inv;
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) Q[i * 3 + j] = W[i * 3] * V[j * 3] + W[i * 3 + 1] * V[j * 3 + 1] + W[i * 3 + 2] * V[j * 3 + 2];
}

// R (float[9]) -> rvec (float[3])
inline void matrix_to_rvec(const float* R9, float* rvec) {
    double A[9], R[9];
    for (int i = 0; i < 9; i++) A[i] = R9[i];
    orthonormalize(A, R);
    double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
    const double s = std::sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
    double c = (R[0] + R[4] + R[8] - 1) * 0.5;
    c = c > 1. ? 1. : (c < -1. ? -1. : c);
    const double theta = std::acos(c);
    if (s < 1e-5) {
        if (c > 0)
            rx = ry = rz = 0;
        else {
            double t;
            t = (R[0] + 1) * 0.5;
            rx = std::sqrt(t > 0. ? t : 0.);
            t = (R[4] + 1) * 0.5;
            ry = std::sqrt(t > 0. ? t : 0.) * (R[1] < 0 ? -1. : 1.);
            t = (R[8] + 1) * 0.5;
            rz = std::sqrt(t > 0. ? t : 0.) * (R[2] < 0 ? -1. : 1.);
            if (std::fabs(rx) < std::fabs(ry) && std::fabs(rx) < std::fabs(rz) && (R[5] > 0) != (ry * rz > 0)) rz = -rz;
            const double n = std::sqrt(rx * rx + ry * ry + rz * rz);
            const double k = theta / n;
            rx *= k, ry *= k, rz *= k;
        }
    } else {
        const double vth = 1 / (2 * s) * theta;
        rx *= vth, ry *= vth, rz *= vth;
    }
    rvec[0] = (float)rx, rvec[1] = (float)ry, rvec[2] = (float)rz;
}

inline double norm3(const float* v) {
    return std::sqrt((double)v[0] * v[0] + (double)v[1] * v[1] + (double)v[2] * v[2]);
}

}  // namespace hm
}  // namespace vb
