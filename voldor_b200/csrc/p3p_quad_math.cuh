// Individually rounded arithmetic for the quad-lane minimal solvers (p3p_twist_quad.cuh, p3p_ap3p_quad.cuh).
//
// The pose hypotheses feed a mode search whose result steers discrete per-pixel decisions, so they have to carry the
// same bits as the reference build's.  That build (nvcc -O3, default -fmad) contracts multiply-adds in two places —
// the front end (visible in its PTX as fma.rn) and ptxas (single-use mul feeding an add/sub; with two such products
// the FIRST operand's product is the one that is fused: tools/ptx_fusion_sites.py, profiles/r02_p3p_fusion_audit.md) —
// and evaluates the double-literal sub-expressions of lambdatwist in FP64.  Every operation below is therefore an
// explicit round-to-nearest primitive: nothing is left to this compiler's contraction choices, and the same source
// gives the same bits on the device (intrinsics) and on the host (IEEE scalar ops under -ffp-contract=off + fmaf),
// which is what lets tests/test_cpu_p3p_quad.py check the solvers against golden hypotheses without a GPU.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define VBQ_FN __host__ __device__ __forceinline__
#else
#define VBQ_FN inline
#endif

namespace vb {
namespace quad {

#if defined(__CUDA_ARCH__)
VBQ_FN float mul(float a, float b) { return __fmul_rn(a, b); }
VBQ_FN float add(float a, float b) { return __fadd_rn(a, b); }
VBQ_FN float sub(float a, float b) { return __fsub_rn(a, b); }
VBQ_FN float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
VBQ_FN float quot(float a, float b) { return __fdiv_rn(a, b); }
VBQ_FN float rcp(float a) { return __frcp_rn(a); }
VBQ_FN float root(float a) { return __fsqrt_rn(a); }
VBQ_FN double dmul(double a, double b) { return __dmul_rn(a, b); }
VBQ_FN double dadd(double a, double b) { return __dadd_rn(a, b); }
VBQ_FN double dsub(double a, double b) { return __dsub_rn(a, b); }
VBQ_FN double dfma(double a, double b, double c) { return __fma_rn(a, b, c); }
VBQ_FN double dquot(double a, double b) { return __ddiv_rn(a, b); }
VBQ_FN double drcp(double a) { return __drcp_rn(a); }
VBQ_FN double droot(double a) { return __dsqrt_rn(a); }
VBQ_FN float narrow(double a) { return __double2float_rn(a); }
#else
// host: compile with -ffp-contract=off (the build of the CPU test does)
VBQ_FN float mul(float a, float b) { return a * b; }
VBQ_FN float add(float a, float b) { return a + b; }
VBQ_FN float sub(float a, float b) { return a - b; }
VBQ_FN float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
VBQ_FN float quot(float a, float b) { return a / b; }
VBQ_FN float rcp(float a) { return 1.0f / a; }
VBQ_FN float root(float a) { return ::sqrtf(a); }
VBQ_FN double dmul(double a, double b) { return a * b; }
VBQ_FN double dadd(double a, double b) { return a + b; }
VBQ_FN double dsub(double a, double b) { return a - b; }
VBQ_FN double dfma(double a, double b, double c) { return ::fma(a, b, c); }
VBQ_FN double dquot(double a, double b) { return a / b; }
VBQ_FN double drcp(double a) { return 1.0 / a; }
VBQ_FN double droot(double a) { return ::sqrt(a); }
VBQ_FN float narrow(double a) { return (float)a; }
#endif
VBQ_FN double widen(float a) { return (double)a; }

// a*b - c*d as the reference build evaluates it: the first product fused, the second rounded on its own
VBQ_FN float diff_of_products(float a, float b, float c, float d) { return fma(a, b, -mul(c, d)); }

struct Vec3f {
    float x, y, z;
};
VBQ_FN Vec3f vec_sub(const Vec3f& a, const Vec3f& b) { return Vec3f{sub(a.x, b.x), sub(a.y, b.y), sub(a.z, b.z)}; }
// accumulate from +0: fma(ax,bx,0) -> fma(ay,by,.) -> fma(az,bz,.)
VBQ_FN float dot_chain(const Vec3f& a, const Vec3f& b) { return fma(a.z, b.z, fma(a.y, b.y, fma(a.x, b.x, 0.f))); }
VBQ_FN Vec3f cross_fused_first(const Vec3f& a, const Vec3f& b) {
    return Vec3f{diff_of_products(a.y, b.z, a.z, b.y), diff_of_products(a.z, b.x, a.x, b.z),
                 diff_of_products(a.x, b.y, a.y, b.x)};
}

struct Pose {
    float R[9];  // row-major
    float t[3];
};

// Sites where the reference's PTX leaves a multi-use product next to an add/sub, so that reading the PTX does not
// settle whether ptxas folded a copy of the product into an FFMA.  Each site is one bit of a per-solver mask; the
// values were fixed by exhaustive search against golden hypotheses of the reference kernels
// (tests/test_cpu_p3p_quad.py on the host, tests/test_gpu_p3p_sites.py on the device: the only assignment reproducing
// every hypothesis bit for bit).  The host test build
// (-DVBQ_SITE_SEARCH) can override the masks at run time to repeat that search.
#if !defined(__CUDACC__) && defined(VBQ_SITE_SEARCH)
extern unsigned vbq_twist_sites, vbq_ap3p_sites;
#define VBQ_TWIST_SITE(bit) ((vbq_twist_sites >> (bit)) & 1u)
#define VBQ_AP3P_SITE(bit) ((vbq_ap3p_sites >> (bit)) & 1u)
#elif defined(VBQ_SITE_SEARCH_DEVICE)  // tests/p3p_device_probe.cu: masks in __constant__ memory declared by the includer
#define VBQ_TWIST_SITE(bit) ((c_twist_sites >> (bit)) & 1u)
#define VBQ_AP3P_SITE(bit) ((c_ap3p_sites >> (bit)) & 1u)
#else
#define VBQ_TWIST_SITE(bit) ((kTwistSitesFused >> (bit)) & 1u)
#define VBQ_AP3P_SITE(bit) ((kAp3pSitesFused >> (bit)) & 1u)
#endif

// c + a*b (sign = +1) or c - a*b (sign = -1) at an undecided site: fused when the site bit is set
VBQ_FN float site_addmul(unsigned fused, float c, float a, float b, bool negate_product) {
    if (fused) return fma(negate_product ? -a : a, b, c);
    return negate_product ? sub(c, mul(a, b)) : add(c, mul(a, b));
}

// squared pixel distance between the projection of X under P and the observation (lambdatwist_p4p.h:30-37, solve_batch_ap3p.cu:362-369)
VBQ_FN float reprojection_error(const Pose& P, const Vec3f& X, float u, float v, float fx, float fy, float cx,
                                float cy) {
    const float px = add(fma(P.R[2], X.z, fma(P.R[0], X.x, mul(P.R[1], X.y))), P.t[0]);
    const float py = add(fma(P.R[5], X.z, fma(P.R[3], X.x, mul(P.R[4], X.y))), P.t[1]);
    const float pz = add(fma(P.R[8], X.z, fma(P.R[6], X.x, mul(P.R[7], X.y))), P.t[2]);
    const float du = sub(add(cx, quot(mul(fx, px), pz)), u);
    const float dv = sub(add(cy, quot(mul(fy, py), pz)), v);
    return fma(du, du, mul(dv, dv));
}

// The reference's scan over the (compacted) candidate list, on the four (exists, error) pairs in slot order: the
// first existing candidate is taken, a later one replaces it only when strictly better (NaN never replaces).
// Returns the winning slot or -1.
VBQ_FN int pick_by_fourth_point(const bool exists[4], const float err[4]) {
    int best = -1;
    float best_err = 0.f;
    for (int q = 0; q < 4; ++q) {
        if (!exists[q]) continue;
        if (best < 0 || best_err > err[q]) best = q, best_err = err[q];
    }
    return best;
}

}  // namespace quad
}  // namespace vb
