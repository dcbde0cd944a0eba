// Lambda-twist P3P (Persson & Nordberg, ECCV 2018) + 4th-point disambiguation, FP32 with the FP64
// sub-expressions the reference build evaluates in double.
//
// Behavioural source: reference lambdatwist/lambdatwist_p4p.h:5-62 (wrapper, reprojection choice),
// lambdatwist_p3p.h:19-294 (solver), solve_cubic.h:15-34 (root2real), :154-210 (cubick),
// solve_eig0.h:11-80 (eigwithknown0), refine_lambda.h:20-102 (Gauss-Newton), instantiated by the reference
// as lambdatwist_p4p<float,float,5> (gpu-kernels/solve_batch_lambdatwist.cu:22-26).  In that instantiation
// every double literal (2.0, 0.5, 1.0, 3.0, 4.0, 1e-4, 1e-10) promotes its sub-expression to FP64 before the
// result is narrowed back to float; those promotions are written out explicitly below as
// D(..) so the rounding points are visible.  Quirks kept: the "+v" branch has no d>0 guard (Q9), the 4th
// point is chosen by smallest squared reprojection with no threshold or cheirality test (Q10).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace vb {
namespace p3p {

#define D(x) ((double)(x))

struct Vec3 {
    float v[3];
    __device__ float& operator[](int i) { return v[i]; }
    __device__ const float& operator[](int i) const { return v[i]; }
};
struct Mat3 {
    float m[9];  // row-major
    __device__ float& operator()(int r, int c) { return m[r * 3 + c]; }
    __device__ const float& operator()(int r, int c) const { return m[r * 3 + c]; }
};

__device__ inline Vec3 make_vec3(float a, float b, float c) {
    Vec3 r;
    r[0] = a, r[1] = b, r[2] = c;
    return r;
}
__device__ inline float dot3(const Vec3& a, const Vec3& b) {
    float sum = 0.f;
    for (int i = 0; i < 3; ++i) sum += a[i] * b[i];
    return sum;
}
__device__ inline Vec3 sub3(const Vec3& a, const Vec3& b) {
    return make_vec3(a[0] - b[0], a[1] - b[1], a[2] - b[2]);
}
__device__ inline Vec3 scale3(const Vec3& a, float s) {
    return make_vec3(a[0] * s, a[1] * s, a[2] * s);
}
__device__ inline Vec3 cross3(const Vec3& a, const Vec3& b) {
    return make_vec3(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
__device__ inline void normalize3(Vec3& a) {
    const float n = sqrtf(dot3(a, a));
    const float si = 1.0f / n;
    for (int i = 0; i < 3; ++i) a[i] *= si;
}
__device__ inline Vec3 mat_vec(const Mat3& A, const Vec3& x) {
    Vec3 r;
    for (int row = 0; row < 3; ++row) {
        float sum = 0.f;
        for (int i = 0; i < 3; ++i) sum += A(row, i) * x[i];
        r[row] = sum;
    }
    return r;
}
__device__ inline Mat3 mat_mat(const Mat3& A, const Mat3& B) {
    Mat3 C;
    for (int row = 0; row < 3; ++row)
        for (int col = 0; col < 3; ++col) {
            float sum = 0.f;
            for (int i = 0; i < 3; ++i) sum += A(row, i) * B(i, col);
            C(row, col) = sum;
        }
    return C;
}
// adjugate / determinant inverse of a 3x3
__device__ inline Mat3 inverse3(const Mat3& a) {
    Mat3 M;
    M(0, 0) = a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1);
    M(0, 1) = a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2);
    M(0, 2) = a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1);
    M(1, 0) = a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2);
    M(1, 1) = a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0);
    M(1, 2) = a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2);
    M(2, 0) = a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0);
    M(2, 1) = a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1);
    M(2, 2) = a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0);
    // a00*M00 + a01*M10 + a02*M20: the reference build keeps the FIRST product as a rounded multiply and fuses
    // the other two (verified in its PTX); written out because the compiler's choice between the two legal
    // contractions of x*y + z*w depends on surrounding code.
    const float idet =
        1.0f / __fmaf_rn(a(0, 2), M(2, 0), __fmaf_rn(a(0, 1), M(1, 0), __fmul_rn(a(0, 0), M(0, 0))));
    for (int i = 0; i < 9; ++i) M.m[i] *= idet;
    return M;
}

// real roots of x^2 + b x + c   (solve_cubic.h:15-34)
__device__ inline bool root2real(float b, float c, float& r1, float& r2) {
    const float v = (float)(D(b * b) - 4.0 * D(c));
    if (v < 0) {
        r1 = r2 = (float)(0.5 * D(b));
        return false;
    }
    const float y = sqrtf(v);
    if (b < 0) {
        r1 = (float)(0.5 * D(-b + y));
        r2 = (float)(0.5 * D(-b - y));
    } else {
        r1 = (float)(2.0 * D(c) / D(-b + y));
        r2 = (float)(2.0 * D(c) / D(-b - y));
    }
    return true;
}

// one well-conditioned real root of r^3 + b r^2 + c r + d by Newton-Raphson from a 2nd-order seed
// (solve_cubic.h:154-210; at most 50 iterations, at least 7, stop at |h| <= 1e-7)
__device__ inline float cubic_root(float b, float c, float d) {
    float r0;
    if (D(b * b) >= 3.0 * D(c)) {
        const float v = (float)sqrt(D(b * b) - 3.0 * D(c));
        const float t1 = (float)(D(-b - v) / (3.0));
        float k = ((t1 + b) * t1 + c) * t1 + d;
        if (D(k) > 0.0) {
            r0 = (float)(D(t1) - sqrt(D(-k) / (3.0 * D(t1) + D(b))));
        } else {
            const float t2 = (float)(D(-b + v) / (3.0));
            k = ((t2 + b) * t2 + c) * t2 + d;
            r0 = (float)(D(t2) + sqrt(D(-k) / (3.0 * D(t2) + D(b))));
        }
    } else {
        r0 = (float)(D(-b) / 3.0);
        if (D(fabsf(((3.0f * r0 + 2.0f * b) * r0 + c))) < 1e-4) r0 += 1;
    }
    for (unsigned int cnt = 0; cnt < 50; ++cnt) {
        const float fx = (((r0 + b) * r0 + c) * r0 + d);
        if ((cnt < 7 || fabsf(fx) > 1e-7f)) {
            const float fpx = ((3.0f * r0 + 2.0f * b) * r0 + c);
            r0 -= fx / fpx;
        } else
            break;
    }
    return r0;
}

// eigen-decomposition of a symmetric 3x3 with one zero eigenvalue (solve_eig0.h:11-80)
__device__ inline void eig_known0(const Mat3& x, Mat3& E, Vec3& L) {
    L[2] = 0;
    Vec3 v3 = make_vec3(x.m[3] * x.m[7] - x.m[6] * x.m[4], x.m[6] * x.m[1] - x.m[7] * x.m[0],
                        x.m[4] * x.m[0] - x.m[3] * x.m[1]);
    normalize3(v3);

    const float x01_squared = x(0, 1) * x(0, 1);
    const float b = -x(0, 0) - x(1, 1) - x(2, 2);
    const float c = -x01_squared - x(0, 2) * x(0, 2) - x(1, 2) * x(1, 2) + x(0, 0) * (x(1, 1) + x(2, 2)) +
                    x(1, 1) * x(2, 2);
    float e1, e2;
    root2real(b, c, e1, e2);
    // NOTE (parity): solve_eig0.h:37-38 orders the pair with `if (|e1| < |e2|) std::swap(e1, e2)`.  std::swap is a
    // host-only function there; in the reference's DEVICE build nvcc drops the call (and with it the whole
    // conditional), so the GPU path never reorders the eigenvalues — verified in the PTX of the reference's
    // solve kernel (no abs/compare between root2real and the eigenvector code).  The reference GPU build is
    // the parity target (north_star), so the pair is deliberately left unordered here.
    L[0] = e1;
    L[1] = e2;

    const float mx0011 = -x(0, 0) * x(1, 1);
    const float prec_0 = x(0, 1) * x(1, 2) - x(0, 2) * x(1, 1);
    const float prec_1 = x(0, 1) * x(0, 2) - x(0, 0) * x(1, 2);

    const float e = e1;
    const float tmp = (float)(1.0 / D(e * (x(0, 0) + x(1, 1)) + mx0011 - e * e + x01_squared));
    float a1 = -(e * x(0, 2) + prec_0) * tmp;
    float a2 = -(e * x(1, 2) + prec_1) * tmp;
    const float rnorm = (float)(D(1.0f) / sqrt(D(a1 * a1 + a2 * a2) + 1.0));
    a1 *= rnorm;
    a2 *= rnorm;

    const float tmp2 = (float)(1.0 / D(e2 * (x(0, 0) + x(1, 1)) + mx0011 - e2 * e2 + x01_squared));
    float a21 = -(e2 * x(0, 2) + prec_0) * tmp2;
    float a22 = -(e2 * x(1, 2) + prec_1) * tmp2;
    const float rnorm2 = (float)(1.0 / sqrt(D(a21 * a21 + a22 * a22) + 1.0));
    a21 *= rnorm2;
    a22 *= rnorm2;

    E.m[0] = a1, E.m[1] = a21, E.m[2] = v3[0];
    E.m[3] = a2, E.m[4] = a22, E.m[5] = v3[1];
    E.m[6] = rnorm, E.m[7] = rnorm2, E.m[8] = v3[2];
}

// Gauss-Newton polish of the three depths (refine_lambda.h:20-102), `iterations` = 5 in the reference
template <int iterations>
__device__ inline void refine_lambda(Vec3& L, float a12, float a13, float a23, float b12, float b13,
                                              float b23) {
    for (int i = 0; i < iterations; ++i) {
        const float l1 = L[0], l2 = L[1], l3 = L[2];
        const float r1 = l1 * l1 + l2 * l2 + b12 * l1 * l2 - a12;
        const float r2 = l1 * l1 + l3 * l3 + b13 * l1 * l3 - a13;
        const float r3 = l2 * l2 + l3 * l3 + b23 * l2 * l3 - a23;
        if (D(fabsf(r1) + fabsf(r2) + fabsf(r3)) < 1e-10) break;

        const float dr1dl1 = (float)((2.0) * D(l1) + D(b12 * l2));
        const float dr1dl2 = (float)((2.0) * D(l2) + D(b12 * l1));
        const float dr2dl1 = (float)((2.0) * D(l1) + D(b13 * l3));
        const float dr2dl3 = (float)((2.0) * D(l3) + D(b13 * l1));
        const float dr3dl2 = (float)((2.0) * D(l2) + D(b23 * l3));
        const float dr3dl3 = (float)((2.0) * D(l3) + D(b23 * l2));

        Vec3 r = make_vec3(r1, r2, r3);
        const float v0 = dr1dl1, v1 = dr1dl2, v3 = dr2dl1, v5 = dr2dl3, v7 = dr3dl2, v8 = dr3dl3;
        const float det = (float)((1.0) / D(-v0 * v5 * v7 - v1 * v3 * v8));
        Mat3 Ji;
        Ji.m[0] = -v5 * v7, Ji.m[1] = -v1 * v8, Ji.m[2] = v1 * v5;
        Ji.m[3] = -v3 * v8, Ji.m[4] = v0 * v8, Ji.m[5] = -v0 * v5;
        Ji.m[6] = v3 * v7, Ji.m[7] = -v0 * v7, Ji.m[8] = -v1 * v3;
        const Vec3 step = scale3(mat_vec(Ji, r), det);
        const Vec3 L1 = sub3(L, step);
        {
            const float m1 = L1[0], m2 = L1[1], m3 = L1[2];
            const float r11 = m1 * m1 + m2 * m2 + b12 * m1 * m2 - a12;
            const float r12 = m1 * m1 + m3 * m3 + b13 * m1 * m3 - a13;
            const float r13 = m2 * m2 + m3 * m3 + b23 * m2 * m3 - a23;
            if (fabsf(r11) + fabsf(r12) + fabsf(r13) > fabsf(r1) + fabsf(r2) + fabsf(r3))
                break;
            else
                L = L1;
        }
    }
}

// candidate depths from one sign of the twist (lambdatwist_p3p.h:153-235)
template <bool GUARD_D>
__device__ inline void twist_branch(float s, const Mat3& V, float a12, float a13, float a23, float b12,
                                             float b13, float b23, Vec3* Ls, int& valid) {
    const float w2 = 1.0f / (s * V.m[1] - V.m[0]);
    const float w0 = (V.m[3] - s * V.m[4]) * w2;
    const float w1 = (V.m[6] - s * V.m[7]) * w2;

    const float a = 1.0f / ((a13 - a12) * w1 * w1 - a12 * b13 * w1 - a12);
    const float b = (a13 * b12 * w1 - a12 * b13 * w0 - 2.0f * w0 * w1 * (a12 - a13)) * a;
    const float c = ((a13 - a12) * w0 * w0 + a13 * b12 * w0 + a13) * a;

    if (D(b * b) - 4.0 * D(c) >= 0) {
        float tau1, tau2;
        root2real(b, c, tau1, tau2);
        const float taus[2] = {tau1, tau2};
        for (int k = 0; k < 2; ++k) {
            if (taus[k] > 0) {
                const float tau = taus[k];
                const float d = a23 / (tau * (b23 + tau) + 1.0f);
                if (!GUARD_D || d > 0) {
                    const float l2 = sqrtf(d);
                    const float l3 = tau * l2;
                    const float l1 = w0 * l2 + w1 * l3;
                    if (l1 >= 0) {
                        Ls[valid] = make_vec3(l1, l2, l3);
                        ++valid;
                    }
                }
            }
        }
    }
}

// y*: unit-less image rays (z = 1 before normalisation), x*: 3-D points.  Returns the number of solutions.
template <int refinement_iterations>
__device__ inline int p3p_solve(Vec3 y1, Vec3 y2, Vec3 y3, const Vec3& x1, const Vec3& x2, const Vec3& x3,
                                         Mat3* Rs, Vec3* Ts) {
    normalize3(y1);
    normalize3(y2);
    normalize3(y3);

    const float b12 = (float)(-2.0 * D(dot3(y1, y2)));
    const float b13 = (float)(-2.0 * D(dot3(y1, y3)));
    const float b23 = (float)(-2.0 * D(dot3(y2, y3)));

    const Vec3 d12 = sub3(x1, x2);
    const Vec3 d13 = sub3(x1, x3);
    const Vec3 d23 = sub3(x2, x3);
    const Vec3 d12xd13 = cross3(d12, d13);

    const float a12 = dot3(d12, d12);
    const float a13 = dot3(d13, d13);
    const float a23 = dot3(d23, d23);

    const float c31 = (float)(-0.5 * D(b13));
    const float c23 = (float)(-0.5 * D(b23));
    const float c12 = (float)(-0.5 * D(b12));
    const float blob = (float)(D(c12 * c23 * c31) - 1.0);

    const float s31_squared = (float)(1.0 - D(c31 * c31));
    const float s23_squared = (float)(1.0 - D(c23 * c23));
    const float s12_squared = (float)(1.0 - D(c12 * c12));

    float p3 = (a13 * (a23 * s31_squared - a13 * s23_squared));
    float p2 = (float)(2.0 * D(blob) * D(a23) * D(a13) + D(a13) * (2.0 * D(a12) + D(a13)) * D(s23_squared) +
                       D(a23 * (a23 - a12) * s31_squared));
    float p1 = (float)(D(a23 * (a13 - a23) * s12_squared - a12 * a12 * s23_squared) -
                       2.0 * D(a12) * D(blob * a23 + a13 * s23_squared));
    float p0 = a12 * (a12 * s23_squared - a23 * s12_squared);

    p3 = (float)(1.0 / D(p3));
    p2 *= p3;
    p1 *= p3;
    p0 *= p3;
    const float g = cubic_root(p2, p1, p0);

    const float A00 = (float)(D(a23) * (1.0 - D(g)));
    const float A01 = (float)(D(a23 * b12) * 0.5);
    const float A02 = (float)(D(a23 * b13 * g) * (-0.5));
    const float A11 = a23 - a12 + a13 * g;
    const float A12 = (float)(D(b23 * (a13 * g - a12)) * 0.5);
    const float A22 = g * (a13 - a23) - a12;
    Mat3 A;
    A.m[0] = A00, A.m[1] = A01, A.m[2] = A02;
    A.m[3] = A01, A.m[4] = A11, A.m[5] = A12;
    A.m[6] = A02, A.m[7] = A12, A.m[8] = A22;

    Mat3 V;
    Vec3 L;
    eig_known0(A, V, L);

    const float v = sqrtf(-L[1] / L[0] > 0 ? -L[1] / L[0] : 0.0f);

    int valid = 0;
    Vec3 Ls[4];
    twist_branch<false>(v, V, a12, a13, a23, b12, b13, b23, Ls, valid);
    twist_branch<true>(-v, V, a12, a13, a23, b12, b13, b23, Ls, valid);

    for (int i = 0; i < valid; ++i) refine_lambda<refinement_iterations>(Ls[i], a12, a13, a23, b12, b13, b23);

    Mat3 X;
    X.m[0] = d12[0], X.m[1] = d13[0], X.m[2] = d12xd13[0];
    X.m[3] = d12[1], X.m[4] = d13[1], X.m[5] = d12xd13[1];
    X.m[6] = d12[2], X.m[7] = d13[2], X.m[8] = d12xd13[2];
    X = inverse3(X);

    for (int i = 0; i < valid; ++i) {
        const Vec3 ry1 = scale3(y1, Ls[i][0]);
        const Vec3 ry2 = scale3(y2, Ls[i][1]);
        const Vec3 ry3 = scale3(y3, Ls[i][2]);
        const Vec3 yd1 = sub3(ry1, ry2);
        const Vec3 yd2 = sub3(ry1, ry3);
        const Vec3 yd1xd2 = cross3(yd1, yd2);
        Mat3 Y;
        Y.m[0] = yd1[0], Y.m[1] = yd2[0], Y.m[2] = yd1xd2[0];
        Y.m[3] = yd1[1], Y.m[4] = yd2[1], Y.m[5] = yd1xd2[1];
        Y.m[6] = yd1[2], Y.m[7] = yd2[2], Y.m[8] = yd1xd2[2];
        Rs[i] = mat_mat(Y, X);
        Ts[i] = sub3(ry1, mat_vec(Rs[i], x1));
    }
    return valid;
}

// P3P on points 1..3, pick the solution that reprojects point 4 best (lambdatwist_p4p.h:5-62).
// y*: pixel coordinates (2 floats), x*: 3-D points (3 floats).
__device__ inline bool p4p_solve(const float* y1, const float* y2, const float* y3, const float* y4,
                                          const float* x1, const float* x2, const float* x3, const float* x4,
                                          float fx, float fy, float cx, float cy, float R[3][3], float t[3]) {
    const Vec3 vy1 = make_vec3((y1[0] - cx) / fx, (y1[1] - cy) / fy, 1.0f);
    const Vec3 vy2 = make_vec3((y2[0] - cx) / fx, (y2[1] - cy) / fy, 1.0f);
    const Vec3 vy3 = make_vec3((y3[0] - cx) / fx, (y3[1] - cy) / fy, 1.0f);
    const Vec3 vx1 = make_vec3(x1[0], x1[1], x1[2]);
    const Vec3 vx2 = make_vec3(x2[0], x2[1], x2[2]);
    const Vec3 vx3 = make_vec3(x3[0], x3[1], x3[2]);

    Mat3 Rs[4];
    Vec3 Ts[4];
    const int n = p3p_solve<5>(vy1, vy2, vy3, vx1, vx2, vx3, Rs, Ts);
    if (n == 0) return false;

    int ns = 0;
    float min_reproj = 0;
    for (int i = 0; i < n; i++) {
        const float X3p = Rs[i](0, 0) * x4[0] + Rs[i](0, 1) * x4[1] + Rs[i](0, 2) * x4[2] + Ts[i][0];
        const float Y3p = Rs[i](1, 0) * x4[0] + Rs[i](1, 1) * x4[1] + Rs[i](1, 2) * x4[2] + Ts[i][1];
        const float Z3p = Rs[i](2, 0) * x4[0] + Rs[i](2, 1) * x4[1] + Rs[i](2, 2) * x4[2] + Ts[i][2];
        const float mu3p = cx + fx * X3p / Z3p;
        const float mv3p = cy + fy * Y3p / Z3p;
        const float reproj = (mu3p - y4[0]) * (mu3p - y4[0]) + (mv3p - y4[1]) * (mv3p - y4[1]);
        if (i == 0 || min_reproj > reproj) {
            ns = i;
            min_reproj = reproj;
        }
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r][c] = Rs[ns](r, c);
    t[0] = Ts[ns][0], t[1] = Ts[ns][1], t[2] = Ts[ns][2];
    return true;
}

#undef D

}  // namespace p3p
}  // namespace vb
