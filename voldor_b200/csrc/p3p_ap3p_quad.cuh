// Algebraic P3P (Ke & Roumeliotis, "An Efficient Algebraic Solution to the Perspective-Three-Point Problem",
// CVPR 2017) laid out for FOUR LANES PER HYPOTHESIS — the alternative minimal solver (`--lambdatwist 0`).
//
// What the reference computes (gpu-kernels/solve_batch_ap3p.cu:152-378, one thread per hypothesis):
//   shared part   two orthonormal frames (k1, nl, k1 x nl) from the 3-D points and (b1, k3, b1 x k3) from the
//                 bearings, the products g1..g7, the quartic in cos(theta1') and its four roots by Ferrari's closed form
//   x4 roots      two Newton steps on each root, then for |root| <= 1 the rotation as a product of three elementary
//                 matrices and the translation; a 4th correspondence picks the pose with the smallest reprojection
// Here the root index IS the lane index inside a quad: every lane evaluates the shared part (identical instructions
// on identical data), polishes and expands only its own root, and the winner of the 4th-point test is found with quad
// shuffles in the reference's scan order.  Arithmetic per root is the reference build's, operation for operation
// (p3p_quad_math.cuh); the transcendental pieces of Ferrari's formula (cbrtf, atan2f, powf, cosf) are the same CUDA
// math-library functions the reference calls, so on the DEVICE hypotheses are bit-identical.  (A host build uses
// libm for those four calls and agrees to rounding only; tests/test_cpu_p3p_quad.py checks it with a tolerance and
// tests/test_gpu_p3p_sites.py checks the device bits.)
//
// Reference behaviour kept because it is visible in results: the complex square root always returns a non-positive
// imaginary part (solve_batch_ap3p.cu:9-15), 0*x terms of the elementary-matrix products are evaluated (they
// propagate NaN/Inf), no cheirality test, smallest squared reprojection wins with a strict comparison.
#pragma once
#include "p3p_quad_math.cuh"

namespace vb {
namespace quad {

// undecided contraction sites (see p3p_quad_math.cuh): a product with two consumers, one bit per consumer
enum Ap3pSite {
    kSiteG3SqInLead = 0,    // c4 = (g5^2 + g1^2) + g3*g3
    kSiteG3SqInC2 = 1,      // c2 = ... - g3*g3
    kSiteG3G4InC3 = 2,      // c3/2 = g4*g3 + (g5 g6 + g1 g2)
    kSiteG3G4InC1 = 3,      // c1/2 = ... - g4*g3
    kSiteG4SqInC2 = 4,      // c2 = g4*g4 + ...
    kSiteG4SqInC0 = 5,      // c0 = ... - g4*g4
    kSiteRatioPlus = 6,     // complex1 + Re(complex2)
    kSiteRatioMinus = 7,    // complex1 - Re(complex2)
    kSiteG1SqInC2 = 8,      // c2 = ... - g1*g1      (the product also feeds an fma as addend)
    kSiteG2SqInC2 = 9       // c2 = g2*g2 + ...      (likewise)
};
// g3*g3, g4*g3 and g4*g4 are folded into both of their consumers; every other site keeps its rounded product
// (search over all 2^10 assignments on the device against 8192 hypotheses of the reference kernel)
constexpr unsigned kAp3pSitesFused = 0x3Fu;

#if defined(__CUDA_ARCH__)
VBQ_FN float lib_cbrt(float a) { return cbrtf(a); }
VBQ_FN float lib_atan2(float y, float x) { return atan2f(y, x); }
VBQ_FN float lib_pow(float a, float b) { return powf(a, b); }
VBQ_FN float lib_cos(float a) { return cosf(a); }
#else
VBQ_FN float lib_cbrt(float a) { return ::cbrtf(a); }
VBQ_FN float lib_atan2(float y, float x) { return ::atan2f(y, x); }
VBQ_FN float lib_pow(float a, float b) { return ::powf(a, b); }
VBQ_FN float lib_cos(float a) { return ::cosf(a); }
#endif

// |a + ib| the way cuComplex.h's cuCabsf evaluates it (scaled by the larger magnitude; inputs already absolute)
VBQ_FN float scaled_hypot_abs(float a, float b) {
    const float v = a > b ? a : b, w = a > b ? b : a;
    const float t = quot(w, v);
    float r = mul(root(fma(t, t, 1.0f)), v);
    if (v == 0.0f || v > 3.402823466e38f || w > 3.402823466e38f) r = add(a, b);
    return r;
}
// real part of the reference's complex square root of (re + i*im) given its modulus: sqrt(|z| (re/|z| + 1) / 2)
VBQ_FN float csqrt_real(float re, float modulus) { return root(mul(mul(modulus, add(quot(re, modulus), 1.0f)), 0.5f)); }
// magnitude of its imaginary part: sqrt(|z| (1 - re/|z|) / 2)   (the reference then forces the sign to be negative)
VBQ_FN float csqrt_imag_abs(float re, float modulus) {
    return fabsf(root(mul(mul(modulus, sub(1.0f, quot(re, modulus))), 0.5f)));
}

struct Ap3pShared {
    Vec3f b1, k3;           // bearing 1, unit normal of (b1, b2)
    float tzx, ntzy, tzz;   // b1 x k3 with the y component stored negated (as the reference build keeps it)
    Vec3f k1, nl;           // unit (X1 - X2), unit (X1 - X3) x k1
    float tmx, ntmy, tmz;   // k1 x nl, y component stored negated
    Vec3f b3p;              // b3 * delta / (k3 . b3)
    float k3b3;
    float g1, g2, g3, g4, g5, g6, g7;
    float c4, c3, c2, c1, c0;  // quartic coefficients, leading first
    float root[4];             // unpolished roots of the quartic
};

// the four roots of c4 x^4 + ... + c0 by Ferrari's method (solve_batch_ap3p.cu:28-82)
VBQ_FN void quartic_roots_ferrari(float c4, float c3, float c2, float c1, float c0, float* r) {
    const float c4sq = mul(c4, c4), c3sq = mul(c3, c3), c4cu = mul(c4, c4sq), c2c4 = mul(c4, c2);
    const float p4 = quot(fma(c2c4, 8.0f, -mul(c3sq, 3.0f)), mul(c4sq, 8.0f));
    const float q4 = quot(fma(c4sq, mul(c1, 8.0f), fma(c3, c3sq, mul(c3, mul(c2c4, -4.0f)))), mul(c4cu, 8.0f));
    const float r4n = fma(c3sq, mul(c2c4, 16.0f),
                          fma(c4sq, mul(c3, mul(c1, -64.0f)), fma(mul(c0, 256.0f), c4cu, -mul(mul(c3sq, c3sq), 3.0f))));
    const float r4 = quot(r4n, mul(mul(c4, c4cu), 256.0f));
    const float p3 = quot(add(r4, quot(mul(p4, p4), 12.0f)), 3.0f);
    const float two_p4 = add(p4, p4);
    const float q3 = quot(fma(q4, mul(q4, -27.0f), fma(p4, mul(r4, 72.0f), -mul(p4, mul(p4, two_p4)))), 432.0f);

    // w = sqrt(q3^2 - p3^3) (complex), then the cube root of -q3 -/+ w
    const float disc = fma(q3, q3, -mul(p3, mul(p3, p3)));
    const float disc_mod = scaled_hypot_abs(fabsf(disc), 0.0f);
    float wx = csqrt_real(disc, disc_mod);
    float wy;  // imaginary part after the reference's sign handling
    if (q3 >= 0.f) {
        wy = csqrt_imag_abs(disc, disc_mod);  // -(-|y|)
        wx = -wx;
    } else {
        const float y0 = -csqrt_imag_abs(disc, disc_mod);
        const float mod2 = scaled_hypot_abs(fabsf(wx), fabsf(y0));
        const float x1 = csqrt_real(wx, mod2);
        wy = -csqrt_imag_abs(wx, mod2);
        wx = x1;
    }
    wx = sub(wx, q3);
    float t;
    if (wy == 0.0f) {
        const float cr = lib_cbrt(wx);
        const float s = add(cr, quot(p3, cr));
        t = add(s, s);
    } else {
        const float theta = lib_atan2(wy, wx);
        const float radius = lib_pow(scaled_hypot_abs(fabsf(wx), fabsf(wy)), 1.0f / 3.0f);
        t = mul(mul(radius, lib_cos(mul(theta, 1.0f / 3.0f))), 4.0f);
    }

    // sqrt(2m) with 2m = t - 2 p4 / 3 (complex square root of a real)
    const float m2 = sub(t, quot(two_p4, 3.0f));
    const float m2_mod = scaled_hypot_abs(fabsf(m2), 0.0f);
    const float sx = csqrt_real(m2, m2_mod);
    const float sy = -csqrt_imag_abs(m2, m2_mod);
    const float shift = quot(-c3, mul(c4, 4.0f));
    const float c1re = add(quot(mul(p4, 4.0f), 3.0f), t);
    // (2 q4 + 0i) / (sx + i sy) with cuComplex.h's scaled division
    const float inv_s = rcp(add(fabsf(sx), fabsf(sy)));
    const float ars = mul(add(q4, q4), inv_s), ais = mul(inv_s, 0.0f), brs = mul(sx, inv_s), bis = mul(inv_s, sy);
    const float inv_n = rcp(fma(brs, brs, mul(bis, bis)));
    const float re_num = fma(ars, brs, mul(ais, bis));
    const float qim = mul(fma(ais, brs, -mul(ars, bis)), inv_n);
    const float half_sx = mul(sx, 0.5f);
    // sqrt(-(complex1 +/- complex2)).real / 2
    float sum_re, dif_re;
    if (VBQ_AP3P_SITE(kSiteRatioPlus))
        sum_re = fma(re_num, inv_n, c1re);
    else
        sum_re = add(c1re, mul(re_num, inv_n));
    if (VBQ_AP3P_SITE(kSiteRatioMinus))
        dif_re = fma(-re_num, inv_n, c1re);
    else
        dif_re = sub(c1re, mul(re_num, inv_n));
    const float sum_im = add(qim, 0.0f), dif_im = sub(0.0f, qim);
    const float mod_p = scaled_hypot_abs(fabsf(-sum_re), fabsf(-sum_im));
    const float half1 = mul(root(mul(mul(mod_p, sub(1.0f, quot(sum_re, mod_p))), 0.5f)), 0.5f);
    const float mod_m = scaled_hypot_abs(fabsf(-dif_re), fabsf(-dif_im));
    const float half2 = mul(root(mul(mul(mod_m, sub(1.0f, quot(dif_re, mod_m))), 0.5f)), 0.5f);
    const float up = add(shift, half_sx), dn = sub(shift, half_sx);
    r[0] = add(up, half1), r[1] = sub(up, half1), r[2] = add(dn, half2), r[3] = sub(dn, half2);
}

VBQ_FN float norm3_mid_first(float x, float y, float z) { return root(fma(z, z, fma(x, x, mul(y, y)))); }
// a.b accumulated the way the reference build does it: the middle product rounded, x and z terms fused around it
VBQ_FN float dot_mid_first(const Vec3f& a, const Vec3f& b) { return fma(a.z, b.z, fma(a.x, b.x, mul(a.y, b.y))); }

VBQ_FN void ap3p_shared(const Vec3f& b1, const Vec3f& b2, const Vec3f& b3, const Vec3f& X1, const Vec3f& X2,
                        const Vec3f& X3, Ap3pShared& S) {
    S.b1 = b1;
    // k1 = (X1 - X2)/|X1 - X2|
    const Vec3f u0 = vec_sub(X1, X2);
    const float nu0 = norm3_mid_first(u0.x, u0.y, u0.z);
    S.k1 = Vec3f{quot(u0.x, nu0), quot(u0.y, nu0), quot(u0.z, nu0)};
    // k3 = (b1 x b2)/|.|, tz = b1 x k3   (y components of the cross products are produced negated)
    const float cx_ = diff_of_products(b1.y, b2.z, b1.z, b2.y), ncy = diff_of_products(b1.x, b2.z, b1.z, b2.x),
                cz_ = diff_of_products(b1.x, b2.y, b1.y, b2.x);
    const float nk3 = root(fma(cz_, cz_, fma(cx_, cx_, mul(ncy, ncy))));
    S.k3 = Vec3f{quot(cx_, nk3), quot(-ncy, nk3), quot(cz_, nk3)};
    S.tzx = diff_of_products(b1.y, S.k3.z, b1.z, S.k3.y);
    S.ntzy = diff_of_products(b1.x, S.k3.z, b1.z, S.k3.x);
    S.tzz = diff_of_products(b1.x, S.k3.y, b1.y, S.k3.x);
    // v1 = b1 x b3, v2 = b2 x b3 (y negated)
    const float v1x = diff_of_products(b1.y, b3.z, b1.z, b3.y), nv1y = diff_of_products(b1.x, b3.z, b1.z, b3.x),
                v1z = diff_of_products(b1.x, b3.y, b1.y, b3.x);
    const float v2x = diff_of_products(b2.y, b3.z, b2.z, b3.y), nv2y = diff_of_products(b2.x, b3.z, b2.z, b3.x),
                v2z = diff_of_products(b2.x, b3.y, b2.y, b3.x);
    const Vec3f u1 = vec_sub(X1, X3);
    const float u1k1 = dot_mid_first(u1, S.k1);
    const float k3b3 = fma(b3.z, S.k3.z, fma(b3.x, S.k3.x, mul(b3.y, S.k3.y)));
    S.k3b3 = k3b3;
    const float f13 = fma(S.k3.z, v1z, fma(S.k3.x, v1x, -mul(nv1y, S.k3.y)));
    const float f15 = mul(k3b3, -u1k1);
    // nl = (u1 x k1)/delta
    const float nx = diff_of_products(u1.y, S.k1.z, u1.z, S.k1.y), nny = diff_of_products(u1.x, S.k1.z, u1.z, S.k1.x),
                nz = diff_of_products(u1.x, S.k1.y, u1.y, S.k1.x);
    const float delta = root(fma(nz, nz, fma(nx, nx, mul(nny, nny))));
    S.nl = Vec3f{quot(nx, delta), quot(-nny, delta), quot(nz, delta)};
    const float f11 = mul(delta, k3b3), f13d = mul(delta, f13);
    const float u2k1 = sub(u1k1, nu0);
    const float f21 = fma(v2z, S.tzz, fma(v2x, S.tzx, mul(nv2y, S.ntzy)));
    const float f22 = mul(nk3, k3b3);
    const float f23 = fma(S.k3.z, v2z, fma(S.k3.x, v2x, -mul(nv2y, S.k3.y)));
    const float f24 = mul(u2k1, f22);
    const float f25 = mul(f21, -u2k1);
    const float f21d = mul(delta, f21), f22d = mul(delta, f22), f23d = mul(delta, f23);
    const float g1 = mul(f13d, f22d);
    const float g2 = fma(f13d, f25, -mul(f15, f23d));
    const float g3 = fma(f11, f23d, -mul(f13d, f21d));
    const float g4 = mul(f24, -f13d);
    const float g5 = mul(f11, f22d);
    const float g6 = fma(f11, f25, -mul(f15, f21d));
    const float g7 = mul(f24, -f15);
    S.g1 = g1, S.g2 = g2, S.g3 = g3, S.g4 = g4, S.g5 = g5, S.g6 = g6, S.g7 = g7;
    // quartic coefficients (solve_batch_ap3p.cu:221-225)
    const float g1sq = mul(g1, g1), g1g2 = mul(g1, g2), g3sq = mul(g3, g3), g3g4 = mul(g4, g3), g2sq = mul(g2, g2),
                g4sq = mul(g4, g4);
    const float lead_a = fma(g5, g5, g1sq);
    S.c4 = VBQ_AP3P_SITE(kSiteG3SqInLead) ? fma(g3, g3, lead_a) : add(lead_a, g3sq);
    const float c3a = fma(g5, g6, g1g2);
    const float c3h = VBQ_AP3P_SITE(kSiteG3G4InC3) ? fma(g4, g3, c3a) : add(g3g4, c3a);
    S.c3 = add(c3h, c3h);
    const float c2a = fma(g7, add(g5, g5), mul(g6, g6));
    const float c2b = VBQ_AP3P_SITE(kSiteG2SqInC2) ? fma(g2, g2, c2a) : add(g2sq, c2a);
    const float c2c = VBQ_AP3P_SITE(kSiteG4SqInC2) ? fma(g4, g4, c2b) : add(g4sq, c2b);
    const float c2d = VBQ_AP3P_SITE(kSiteG1SqInC2) ? fma(-g1, g1, c2c) : sub(c2c, g1sq);
    S.c2 = VBQ_AP3P_SITE(kSiteG3SqInC2) ? fma(-g3, g3, c2d) : sub(c2d, g3sq);
    const float c1a = fma(g7, g6, -g1g2);
    const float c1h = VBQ_AP3P_SITE(kSiteG3G4InC1) ? fma(-g4, g3, c1a) : sub(c1a, g3g4);
    S.c1 = add(c1h, c1h);
    const float c0a = fma(g7, g7, -g2sq);
    S.c0 = VBQ_AP3P_SITE(kSiteG4SqInC0) ? fma(-g4, g4, c0a) : sub(c0a, g4sq);
    quartic_roots_ferrari(S.c4, S.c3, S.c2, S.c1, S.c0, S.root);
    // third axis of the point frame and the scaled third bearing
    S.tmx = diff_of_products(S.k1.y, S.nl.z, S.k1.z, S.nl.y);
    S.ntmy = diff_of_products(S.k1.x, S.nl.z, S.k1.z, S.nl.x);
    S.tmz = diff_of_products(S.k1.x, S.nl.y, S.k1.y, S.nl.x);
    const float scale = quot(delta, k3b3);
    S.b3p = Vec3f{mul(b3.x, scale), mul(b3.y, scale), mul(b3.z, scale)};
}

// two Newton steps on one root (solve_batch_ap3p.cu:85-98)
VBQ_FN float quartic_polish(const Ap3pShared& S, float x) {
    const float four_c4 = mul(S.c4, 4.0f), three_c3 = mul(S.c3, 3.0f), two_c2 = add(S.c2, S.c2);
    for (int it = 0; it < 2; ++it) {
        const float err = fma(x, fma(x, fma(x, fma(S.c4, x, S.c3), S.c2), S.c1), S.c0);
        const float der = fma(x, fma(x, fma(four_c4, x, three_c3), two_c2), S.c1);
        x = sub(x, quot(err, der));
    }
    return x;
}

// pose of root `ct` = cos(theta1'); false when |ct| > 1 (solve_batch_ap3p.cu:249-289)
VBQ_FN bool ap3p_pose(const Ap3pShared& S, const Vec3f& X3, float ct, Pose& P) {
    if (fabsf(ct) > 1.0f) return false;
    float st = root(fma(-ct, ct, 1.0f));
    st = (S.k3b3 > 0.f) ? st : -st;
    const float scale = quot(st, fma(ct, fma(S.g5, ct, S.g6), S.g7));
    const float c3 = mul(fma(S.g1, ct, S.g2), scale), s3 = mul(fma(S.g3, ct, S.g4), scale);
    const float ss = mul(st, s3), sc = mul(st, c3), cs = mul(ct, s3), cc = mul(ct, c3);
    // M = [k1 | nl | k1 x nl] * C13, row by row (tm_y and tz_y enter with their stored negations)
    const float k1c[3] = {S.k1.x, S.k1.y, S.k1.z}, nlc[3] = {S.nl.x, S.nl.y, S.nl.z};
    float M0[3], M1[3], M2[3];
    M0[0] = fma(S.tmx, cs, fma(k1c[0], c3, mul(nlc[0], ss)));
    M0[1] = fma(-S.ntmy, cs, fma(k1c[1], c3, mul(nlc[1], ss)));
    M0[2] = fma(S.tmz, cs, fma(k1c[2], c3, mul(nlc[2], ss)));
    M1[0] = fma(-S.tmx, st, fma(k1c[0], 0.0f, mul(nlc[0], ct)));
    M1[1] = fma(S.ntmy, st, fma(k1c[1], 0.0f, mul(nlc[1], ct)));
    M1[2] = fma(-S.tmz, st, fma(k1c[2], 0.0f, mul(nlc[2], ct)));
    M2[0] = fma(S.tmx, cc, fma(nlc[0], sc, -mul(k1c[0], s3)));
    M2[1] = fma(-S.ntmy, cc, fma(nlc[1], sc, -mul(k1c[1], s3)));
    M2[2] = fma(S.tmz, cc, fma(nlc[2], sc, -mul(k1c[2], s3)));
    // Q = M * [b1; k3; tz] (rows), the camera rotation is its transpose
    float Q[9];
    for (int i = 0; i < 3; ++i) {
        Q[i * 3 + 0] = fma(S.tzx, M2[i], fma(S.k3.x, M1[i], mul(S.b1.x, M0[i])));
        Q[i * 3 + 1] = fma(-S.ntzy, M2[i], fma(S.k3.y, M1[i], mul(S.b1.y, M0[i])));
        Q[i * 3 + 2] = fma(S.tzz, M2[i], fma(S.k3.z, M1[i], mul(S.b1.z, M0[i])));
    }
    const float b3pc[3] = {S.b3p.x, S.b3p.y, S.b3p.z};
    for (int j = 0; j < 3; ++j) {
        const float back = fma(X3.z, Q[6 + j], fma(X3.x, Q[j], mul(X3.y, Q[3 + j])));
        P.t[j] = fma(b3pc[j], st, -back);
        P.R[j * 3 + 0] = Q[j], P.R[j * 3 + 1] = Q[3 + j], P.R[j * 3 + 2] = Q[6 + j];
    }
    return true;
}

// unit bearing as solve_all() builds it (solve_batch_ap3p.cu:298-324): squares accumulated y first
VBQ_FN Vec3f unit_bearing_ap3p(float u, float v, float fx, float fy, float cx, float cy) {
    const float mx = quot(sub(u, cx), fx), my = quot(sub(v, cy), fy);
    const float s = rcp(root(add(fma(mx, mx, mul(my, my)), 1.0f)));
    return Vec3f{mul(mx, s), mul(my, s), s};
}

// everything lane `slot` does for one hypothesis
VBQ_FN bool ap3p_lane(int slot, const float* uv /*[4][2]*/, const Vec3f* X /*[4]*/, float fx, float fy, float cx,
                      float cy, Pose& P, float& err) {
    Ap3pShared S;
    ap3p_shared(unit_bearing_ap3p(uv[0], uv[1], fx, fy, cx, cy), unit_bearing_ap3p(uv[2], uv[3], fx, fy, cx, cy),
                unit_bearing_ap3p(uv[4], uv[5], fx, fy, cx, cy), X[0], X[1], X[2], S);
    const float ct = quartic_polish(S, S.root[slot]);
    if (!ap3p_pose(S, X[2], ct, P)) return false;
    err = reprojection_error(P, X[3], uv[6], uv[7], fx, fy, cx, cy);
    return true;
}

}  // namespace quad
}  // namespace vb
