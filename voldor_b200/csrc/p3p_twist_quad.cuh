// Lambda-twist P3P (Persson & Nordberg, "Lambda Twist: An Accurate Fast Robust Perspective Three Point Solver",
// ECCV 2018) laid out for FOUR LANES PER HYPOTHESIS.
//
// What the reference computes (gpu-kernels/solve_batch_lambdatwist.cu:11-42 -> lambdatwist/lambdatwist_p4p.h:5-62,
// lambdatwist_p3p.h:19-294, solve_cubic.h, solve_eig0.h, refine_lambda.h; one thread walks everything serially):
//   shared part   unit bearings, the three cosines and squared side lengths, one root g of the cubic that makes
//                 D1 + g*D2 degenerate, the two non-trivial eigen-pairs of that degenerate quadric
//   x2 signs      the degenerate quadric splits into two planes (+sigma / -sigma); each gives a line w0,w1 and a
//                 quadratic in tau = lambda3/lambda2
//   x2 roots      every positive tau gives one depth triple (lambda1, lambda2, lambda3)
//   per triple    <= 5 Gauss-Newton steps on the three distance constraints, the rigid pose from the two triangles,
//                 reprojection of a 4th point; the pose with the smallest reprojection error wins
// The decomposition here: the (sign, root) pair IS the lane index inside a quad.  All four lanes evaluate the shared
// part (same instructions, same data: free in SIMT), then lane q = 2*sign + root works on its own candidate only, so
// the dependent chain a hypothesis pays for is  shared + 1 x (plane, tau, polish, pose, reprojection)  instead of
// shared + 2 planes + 4 x (...); the winner is picked with quad shuffles in the reference's scan order (slot 0..3,
// strictly-smaller replaces).  The per-candidate arithmetic is the reference build's, operation for operation
// (p3p_quad_math.cuh explains how the rounding points were established), so hypotheses are bit-identical.
//
// Quirks of the reference that are visible in results and therefore kept: sub-expressions with
// double literals run in FP64; the "+sigma" plane accepts a candidate without checking d > 0 (a negative d dies at the
// lambda1 >= 0 test through NaN), the "-sigma" plane checks it; the eigenvalue pair is not ordered by magnitude in the
// device build; no cheirality test and no reprojection threshold when choosing by the 4th point.
#pragma once
#include "p3p_quad_math.cuh"

namespace vb {
namespace quad {

// undecided contraction sites of this solver (see p3p_quad_math.cuh)
enum TwistSite {
    kSiteA11 = 0,        // D23 - D12 + D13*g
    kSiteA12 = 1,        // D13*g - D12
    kSiteW0PlusA = 2,    // V10 - V11*sigma : fold V11*sigma
    kSiteW0PlusB = 3,    //                   fold the product that makes V10
    kSiteW1Plus = 4,     // V20 - sigma*V21
    kSiteW0MinusA = 5,   // V10 + V11*sigma : fold V11*sigma
    kSiteW0MinusB = 6,   //                   fold the product that makes V10
    kSiteW1Minus = 7,    // sigma*V21 + V20
    kSiteTranslation = 8 // lambda1*b1 - R*X1
};
// the two products D13*g are folded into their consumers, every other site keeps its rounded product
// (search over all 2^9 assignments against 1024 golden hypotheses: this one alone matches all of them)
constexpr unsigned kTwistSitesFused = (1u << kSiteA11) | (1u << kSiteA12);

// everything the four candidates of one hypothesis have in common
struct TwistShared {
    Vec3f b1, b2, b3;     // unit bearings
    float g12, g13, g23;  // -2 * cos(angle between bearings)
    float D12, D13, D23;  // squared side lengths of the 3-D triangle
    Vec3f e12, e13;       // X1 - X2, X1 - X3
    float V00, V10, V20;  // eigenvector of the first non-zero eigenvalue (third component = its normaliser)
    float V01, V11, V21;  // ... of the second
    float a2v, rn1;       // the two factors of V10 (undecided site kSiteW0*B)
    float sigma;          // sqrt(max(-e2/e1, 0))
};

VBQ_FN Vec3f unit_bearing(float u, float v, float fx, float fy, float cx, float cy) {
    const float mx = quot(sub(u, cx), fx), my = quot(sub(v, cy), fy);
    const float s = rcp(root(add(fma(my, my, fma(mx, mx, 0.f)), 1.0f)));
    return Vec3f{mul(s, mx), mul(s, my), s};
}

// one real root of x^3 + c2 x^2 + c1 x + c0, Newton-Raphson from a stationary-point seed (>= 7, <= 50 steps)
VBQ_FN float monic_cubic_root(float c2, float c1, float c0) {
    float r;
    const double c2sq = widen(mul(c2, c2)), three_c1 = dmul(widen(c1), 3.0);
    if (!(three_c1 > c2sq)) {  // two stationary points: start beyond the one on the root's side
        const float v = narrow(droot(dsub(c2sq, three_c1)));
        const float t1 = quot(sub(-c2, v), 3.0f);
        const float k1 = fma(t1, fma(t1, add(c2, t1), c1), c0);
        if (k1 > 0.f) {
            const double dt = widen(t1);
            r = narrow(dsub(dt, droot(dquot(widen(-k1), dfma(dt, 3.0, widen(c2))))));
        } else {
            const float t2 = quot(sub(v, c2), 3.0f);
            const float k2 = fma(t2, fma(t2, add(c2, t2), c1), c0);
            const double dt = widen(t2);
            r = narrow(dadd(droot(dquot(widen(-k2), dfma(dt, 3.0, widen(c2)))), dt));
        }
    } else {  // monotone cubic: start at the inflection point, nudged off a flat slope
        const float ti = quot(c2, -3.0f);
        const float slope = fma(ti, fma(c2, 2.0f, mul(ti, 3.0f)), c1);
        r = (widen(fabsf(slope)) < 1e-4) ? add(ti, 1.0f) : ti;
    }
    const float two_c2 = add(c2, c2);
    for (int step = 0; step < 50; ++step) {
        const float f = fma(r, fma(r, add(c2, r), c1), c0);
        if (step >= 7 && !(fabsf(f) > 1e-7f)) break;
        const float df = fma(r, fma(r, 3.0f, two_c2), c1);
        r = sub(r, quot(f, df));
    }
    return r;
}

// roots of x^2 + b x + c given disc = b^2 - 4c >= 0 (or NaN) and twice_c = 2c in double
VBQ_FN void monic_quadratic_roots(float b, float disc, double twice_c, float& r1, float& r2) {
    const float y = root(disc);
    if (b < 0.f) {
        r1 = mul(sub(y, b), 0.5f);
        r2 = mul(sub(-b, y), 0.5f);
    } else {
        r1 = narrow(dquot(twice_c, widen(sub(y, b))));
        r2 = narrow(dquot(twice_c, widen(sub(-b, y))));
    }
}

VBQ_FN void twist_shared(const Vec3f& b1, const Vec3f& b2, const Vec3f& b3, const Vec3f& X1, const Vec3f& X2,
                         const Vec3f& X3, TwistShared& S) {
    S.b1 = b1, S.b2 = b2, S.b3 = b3;
    S.g12 = mul(dot_chain(b1, b2), -2.0f);
    S.g13 = mul(dot_chain(b1, b3), -2.0f);
    S.g23 = mul(dot_chain(b2, b3), -2.0f);
    S.e12 = vec_sub(X1, X2), S.e13 = vec_sub(X1, X3);
    const Vec3f e23 = vec_sub(X2, X3);
    const float D12 = dot_chain(S.e12, S.e12), D13 = dot_chain(S.e13, S.e13), D23 = dot_chain(e23, e23);
    S.D12 = D12, S.D13 = D13, S.D23 = D23;

    // cubic in g: det(D1 + g*D2) = 0, coefficients k3..k0 (lambdatwist_p3p.h:58-81)
    const float h13 = mul(S.g13, -0.5f), h23 = mul(S.g23, -0.5f), h12 = mul(S.g12, -0.5f);  // cosines
    const float mix = fma(h13, mul(h12, h23), -1.0f);
    const float S13 = fma(-h13, h13, 1.0f), S23 = fma(-h23, h23, 1.0f), S12 = fma(-h12, h12, 1.0f);  // sines squared
    const float S23D13 = mul(S23, D13);
    const float k3 = mul(D13, fma(S13, D23, -S23D13));
    const double dD23 = widen(D23), dD13 = widen(D13), two_D12 = dadd(widen(D12), widen(D12));
    const double k2a = dmul(dmul(dadd(widen(mix), widen(mix)), dD23), dD13);
    const double k2b = dfma(dmul(dadd(two_D12, dD13), dD13), widen(S23), k2a);
    const float D23mD12 = sub(D23, D12), D13mD23 = sub(D13, D23);
    const float k2 = narrow(dadd(k2b, widen(mul(S13, mul(D23, D23mD12)))));
    const float k1a = fma(S12, mul(D23, D13mD23), -mul(S23, mul(D12, D12)));
    const float k1 = narrow(dfma(-two_D12, widen(fma(mix, D23, S23D13)), widen(k1a)));
    const float k0 = mul(D12, fma(S23, D12, -mul(S12, D23)));
    const float ik3 = rcp(k3);
    const float g = monic_cubic_root(mul(ik3, k2), mul(ik3, k1), mul(k0, ik3));

    // the degenerate quadric A = D1 + g*D2 (symmetric, one zero eigenvalue) (lambdatwist_p3p.h:95-117)
    const float A00 = narrow(dmul(dsub(1.0, widen(g)), dD23));
    const float A01 = mul(mul(S.g12, D23), 0.5f);
    const float A02 = mul(mul(mul(S.g13, D23), g), -0.5f);
    const float A11 = site_addmul(VBQ_TWIST_SITE(kSiteA11), D23mD12, D13, g, false);
    const float A12 = mul(mul(S.g23, site_addmul(VBQ_TWIST_SITE(kSiteA12), -D12, D13, g, false)), 0.5f);
    const float A22 = fma(D13mD23, g, -D12);

    // its two non-zero eigenvalues: roots of x^2 + qb x + qc (solve_eig0.h:27-36)
    const float A01sq = mul(A01, A01);
    const float nA00 = -A00;
    const float qb = sub(sub(nA00, A11), A22);
    const float qc = fma(A11, A22, fma(add(A11, A22), A00, fma(-A12, A12, fma(-A02, A02, -A01sq))));
    const double dqc = widen(qc);
    const float qdisc = narrow(dfma(dqc, -4.0, widen(mul(qb, qb))));
    float ev1, ev2;
    if (qdisc < 0.f) {
        ev1 = ev2 = mul(qb, 0.5f);
    } else {
        monic_quadratic_roots(qb, qdisc, dadd(dqc, dqc), ev1, ev2);
    }

    // eigenvectors (a1, a2, 1)/norm of each eigenvalue (solve_eig0.h:45-73)
    const float m0 = diff_of_products(A01, A12, A02, A11), m1 = diff_of_products(A01, A02, A12, A00);
    const float trace2 = add(A11, A00);
    float comp1[2], comp2[2], norm[2];
    for (int k = 0; k < 2; ++k) {
        const float ev = k == 0 ? ev1 : ev2;
        const float den = add(A01sq, fma(-ev, ev, fma(A11, nA00, mul(trace2, ev))));
        const float iden = rcp(den);
        const float a1 = mul(iden, -fma(A02, ev, m0)), a2 = mul(iden, -fma(A12, ev, m1));
        const float rn = narrow(drcp(droot(dadd(widen(fma(a1, a1, mul(a2, a2))), 1.0))));
        comp1[k] = mul(a1, rn), comp2[k] = a2, norm[k] = rn;
    }
    S.V00 = comp1[0], S.a2v = comp2[0], S.rn1 = norm[0];
    S.V10 = mul(comp2[0], norm[0]), S.V20 = norm[0];
    S.V01 = comp1[1], S.V11 = mul(comp2[1], norm[1]), S.V21 = norm[1];
    S.sigma = root(fmaxf(quot(-ev2, ev1), 0.f));
}

// candidate `slot` = 2*plane + root of hypothesis state S: depth triple (l1,l2,l3); false if it does not exist
VBQ_FN bool twist_candidate(const TwistShared& S, int slot, float& l1, float& l2, float& l3) {
    const bool minus = (slot >> 1) != 0;
    const float sg = minus ? -S.sigma : S.sigma;
    // line of the plane: lambda1 = w0*lambda2 + w1*lambda3 (lambdatwist_p3p.h:153-160, 195-203)
    const float w2 = rcp(fma(S.V01, sg, -S.V00));
    const float V11s = mul(S.V11, S.sigma), V21s = mul(S.sigma, S.V21);
    float w0n, w1n;
    if (!minus) {
        if (VBQ_TWIST_SITE(kSiteW0PlusA))
            w0n = fma(-S.V11, S.sigma, S.V10);
        else if (VBQ_TWIST_SITE(kSiteW0PlusB))
            w0n = fma(S.a2v, S.rn1, -V11s);
        else
            w0n = sub(S.V10, V11s);
        w1n = site_addmul(VBQ_TWIST_SITE(kSiteW1Plus), S.V20, S.sigma, S.V21, true);
    } else {
        if (VBQ_TWIST_SITE(kSiteW0MinusA))
            w0n = fma(S.V11, S.sigma, S.V10);
        else if (VBQ_TWIST_SITE(kSiteW0MinusB))
            w0n = fma(S.a2v, S.rn1, V11s);
        else
            w0n = add(S.V10, V11s);
        w1n = VBQ_TWIST_SITE(kSiteW1Minus) ? fma(S.sigma, S.V21, S.V20) : add(V21s, S.V20);
    }
    const float w0 = mul(w0n, w2), w1 = mul(w1n, w2);
    // quadratic tau^2 + qb tau + qc = 0 in tau = lambda3 / lambda2 (lambdatwist_p3p.h:162-166)
    const float dD = sub(S.D13, S.D12);
    const float pD12 = mul(S.g13, S.D12), pD13 = mul(S.g12, S.D13);
    const float lead = rcp(sub(fma(w1, mul(dD, w1), -mul(pD12, w1)), S.D12));
    const float mid = fma(-sub(S.D12, S.D13), mul(w1, add(w0, w0)), fma(pD13, w1, -mul(pD12, w0)));
    const float qb = mul(mid, lead);
    const float qc = mul(add(S.D13, fma(pD13, w0, mul(w0, mul(dD, w0)))), lead);
    const double dqc = widen(qc);
    const double ddisc = dfma(dqc, -4.0, widen(mul(qb, qb)));
    if (!(ddisc >= 0.0)) return false;
    float t1, t2;
    monic_quadratic_roots(qb, narrow(ddisc), dadd(dqc, dqc), t1, t2);
    const float tau = (slot & 1) ? t2 : t1;
    if (!(tau > 0.f)) return false;
    const float d = quot(S.D23, fma(tau, add(S.g23, tau), 1.0f));
    if (minus && !(d > 0.f)) return false;  // only this plane checks d (reference quirk)
    l2 = root(d);
    l3 = mul(tau, l2);
    l1 = fma(w0, l2, mul(w1, l3));
    return !(l1 < 0.f) && l1 == l1;
}

// residuals of the three distance constraints and their absolute sum (refine_lambda.h:26-33)
VBQ_FN float twist_residuals(const TwistShared& S, float l1, float l2, float l3, float& r1, float& r2, float& r3,
                             float& p12, float& p13, float& p23) {
    const float l3sq = mul(l3, l3);
    p12 = mul(S.g12, l1), p13 = mul(S.g13, l1), p23 = mul(S.g23, l2);
    r1 = sub(fma(p12, l2, fma(l1, l1, mul(l2, l2))), S.D12);
    r2 = sub(fma(p13, l3, fma(l1, l1, l3sq)), S.D13);
    r3 = sub(fma(p23, l3, fma(l2, l2, l3sq)), S.D23);
    return add(fabsf(r3), add(fabsf(r1), fabsf(r2)));
}

// <= 5 Gauss-Newton steps; a step is only taken when it does not increase the residual (refine_lambda.h:20-102)
VBQ_FN void twist_polish(const TwistShared& S, float& l1, float& l2, float& l3) {
    for (int it = 0; it < 5; ++it) {
        float r1, r2, r3, p12, p13, p23;
        const float res = twist_residuals(S, l1, l2, l3, r1, r2, r3, p12, p13, p23);
        if (widen(res) < 1e-10) break;
        const double d1 = widen(l1), d2 = widen(l2), d3 = widen(l3);
        // Jacobian entries (2*li + g*lj evaluated in double, narrowed)
        const float j11 = narrow(dfma(d1, 2.0, widen(mul(S.g12, l2)))), j12 = narrow(dfma(d2, 2.0, widen(p12)));
        const float j21 = narrow(dfma(d1, 2.0, widen(mul(S.g13, l3)))), j23 = narrow(dfma(d3, 2.0, widen(p13)));
        const float j32 = narrow(dfma(d2, 2.0, widen(mul(S.g23, l3)))), j33 = narrow(dfma(d3, 2.0, widen(p23)));
        const float nj11j23 = mul(-j11, j23), j12j21 = mul(j12, j21);
        const float idet = rcp(fma(nj11j23, j32, -mul(j12j21, j33)));
        const float s1 = fma(mul(j12, j23), r3, fma(-mul(j12, j33), r2, fma(-mul(j23, j32), r1, 0.f)));
        const float s2 = fma(nj11j23, r3, fma(mul(j11, j33), r2, fma(-mul(j21, j33), r1, 0.f)));
        const float s3 = fma(-j12j21, r3, fma(-mul(j11, j32), r2, fma(mul(j21, j32), r1, 0.f)));
        const float n1 = fma(-idet, s1, l1), n2 = fma(-idet, s2, l2), n3 = fma(-idet, s3, l3);
        float q1, q2, q3, u12, u13, u23;
        const float res_new = twist_residuals(S, n1, n2, n3, q1, q2, q3, u12, u13, u23);
        if (res_new > res) break;
        l1 = n1, l2 = n2, l3 = n3;
    }
}

// inverse of [e12 | e13 | e12 x e13] by the adjugate (lambdatwist_p3p.h:256-262, matrix.h inverse of a 3x3)
struct TriangleFrameInverse {
    float r0[3], r1[3], r2[3];
};
VBQ_FN void triangle_frame_inverse(const TwistShared& S, TriangleFrameInverse& Xi) {
    const Vec3f a = S.e12, b = S.e13;
    const Vec3f c = cross_fused_first(a, b);
    const float M00 = diff_of_products(c.z, b.y, b.z, c.y), M01 = diff_of_products(b.z, c.x, b.x, c.z),
                M02 = diff_of_products(b.x, c.y, c.x, b.y);
    const float M10 = diff_of_products(c.y, a.z, a.y, c.z), M11 = diff_of_products(c.z, a.x, c.x, a.z),
                M12 = diff_of_products(c.x, a.y, a.x, c.y);
    const float idet = rcp(fma(c.x, c.x, fma(M10, b.x, mul(M00, a.x))));
    Xi.r0[0] = mul(idet, M00), Xi.r0[1] = mul(idet, M01), Xi.r0[2] = mul(idet, M02);
    Xi.r1[0] = mul(idet, M10), Xi.r1[1] = mul(idet, M11), Xi.r1[2] = mul(idet, M12);
    Xi.r2[0] = mul(idet, c.x), Xi.r2[1] = mul(idet, c.y), Xi.r2[2] = mul(idet, c.z);
}

// rigid pose that maps the 3-D triangle onto the depth-scaled bearings (lambdatwist_p3p.h:264-290)
VBQ_FN void twist_pose(const TwistShared& S, const TriangleFrameInverse& Xi, const Vec3f& X1, float l1, float l2,
                       float l3, Pose& P) {
    const Vec3f y1 = Vec3f{mul(l1, S.b1.x), mul(l1, S.b1.y), mul(l1, S.b1.z)};
    const Vec3f u = Vec3f{fma(-l2, S.b2.x, y1.x), fma(-l2, S.b2.y, y1.y), fma(-l2, S.b2.z, y1.z)};
    const Vec3f v = Vec3f{fma(-l3, S.b3.x, y1.x), fma(-l3, S.b3.y, y1.y), fma(-l3, S.b3.z, y1.z)};
    const Vec3f n = cross_fused_first(u, v);
    const float ucomp[3] = {u.x, u.y, u.z}, vcomp[3] = {v.x, v.y, v.z}, ncomp[3] = {n.x, n.y, n.z};
    const float y1c[3] = {y1.x, y1.y, y1.z}, b1c[3] = {S.b1.x, S.b1.y, S.b1.z};
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c)
            P.R[r * 3 + c] = fma(ncomp[r], Xi.r2[c], fma(vcomp[r], Xi.r1[c], fma(ucomp[r], Xi.r0[c], 0.f)));
        const float RX = fma(P.R[r * 3 + 2], X1.z, fma(P.R[r * 3 + 1], X1.y, fma(P.R[r * 3 + 0], X1.x, 0.f)));
        P.t[r] = VBQ_TWIST_SITE(kSiteTranslation) ? fma(l1, b1c[r], -RX) : sub(y1c[r], RX);
    }
}

// everything lane `slot` does for one hypothesis: returns whether its candidate exists; pose + error if it does
VBQ_FN bool twist_lane(int slot, const float* uv /*[4][2]*/, const Vec3f* X /*[4]*/, float fx, float fy, float cx,
                       float cy, Pose& P, float& err) {
    TwistShared S;
    twist_shared(unit_bearing(uv[0], uv[1], fx, fy, cx, cy), unit_bearing(uv[2], uv[3], fx, fy, cx, cy),
                 unit_bearing(uv[4], uv[5], fx, fy, cx, cy), X[0], X[1], X[2], S);
    float l1, l2, l3;
    if (!twist_candidate(S, slot, l1, l2, l3)) return false;
    twist_polish(S, l1, l2, l3);
    TriangleFrameInverse Xi;
    triangle_frame_inverse(S, Xi);
    twist_pose(S, Xi, X[0], l1, l2, l3, P);
    err = reprojection_error(P, X[3], uv[6], uv[7], fx, fy, cx, cy);
    return true;
}

}  // namespace quad
}  // namespace vb
