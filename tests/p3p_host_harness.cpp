// TEST INFRASTRUCTURE.  Host build of the product's quad-lane minimal solvers (voldor_b200/csrc/p3p_*_quad.cuh are
// written with individually rounded operations, so the host evaluates exactly what the device evaluates): lets
// tests/test_cpu_p3p_quad.py compare them with golden hypotheses of the reference kernels without a GPU, and search the
// undecided contraction sites.  Built by the test with  g++ -O2 -ffp-contract=off -DVBQ_SITE_SEARCH.
#define QUALIFIERS static inline
#include <cuda_runtime.h>
#include <curand_kernel.h>

#include <cmath>
#include <cstring>

#include "../voldor_b200/csrc/p3p_twist_quad.cuh"
#ifdef HAVE_AP3P_QUAD
#include "../voldor_b200/csrc/p3p_ap3p_quad.cuh"
#endif

namespace vb {
namespace quad {
unsigned vbq_twist_sites = 0, vbq_ap3p_sites = 0;
}
}  // namespace vb

using namespace vb::quad;

extern "C" {

void harness_set_sites(int solver, unsigned mask) { (solver == 0 ? vbq_twist_sites : vbq_ap3p_sites) = mask; }
unsigned harness_default_sites(int solver) {
#ifdef HAVE_AP3P_QUAD
    if (solver == 1) return kAp3pSitesFused;
#endif
    return solver == 0 ? kTwistSitesFused : 0u;
}

// the reference sampler's indices: int(curand_uniform * n_pts), 4 per hypothesis (solve_batch_lambdatwist.cu:16-19)
void harness_indices(int n_poses, int n_pts, int* idx4) {
    for (int h = 0; h < n_poses; h++) {
        curandStateXORWOW_t st;
        curand_init(233ULL, (unsigned long long)h, 0ULL, &st);
        for (int k = 0; k < 4; k++) idx4[h * 4 + k] = (int)(curand_uniform(&st) * (float)n_pts);
    }
}

// scalar driver: the four lanes of a quad one after the other, then the selection scan.
// solver: 0 = lambda-twist, 1 = AP3P.  Outputs R (9) and t (3) per hypothesis, NaN when no candidate exists.
int harness_solve(int solver, const float* p2s, const float* p3s, const int* idx4, int n_poses, float fx, float fy,
                  float cx, float cy, float* R_out, float* t_out, int* slot_out) {
    for (int h = 0; h < n_poses; h++) {
        float uv[8];
        Vec3f X[4];
        for (int k = 0; k < 4; k++) {
            const int i = idx4[h * 4 + k];
            uv[2 * k] = p2s[2 * i], uv[2 * k + 1] = p2s[2 * i + 1];
            X[k] = Vec3f{p3s[3 * i], p3s[3 * i + 1], p3s[3 * i + 2]};
        }
        Pose P[4];
        bool exists[4];
        float err[4];
        for (int q = 0; q < 4; q++) {
            err[q] = 0.f;
#ifdef HAVE_AP3P_QUAD
            exists[q] = solver == 0 ? twist_lane(q, uv, X, fx, fy, cx, cy, P[q], err[q])
                                    : ap3p_lane(q, uv, X, fx, fy, cx, cy, P[q], err[q]);
#else
            if (solver != 0) return 1;
            exists[q] = twist_lane(q, uv, X, fx, fy, cx, cy, P[q], err[q]);
#endif
        }
        const int best = pick_by_fourth_point(exists, err);
        if (slot_out) slot_out[h] = best;
        for (int k = 0; k < 9; k++) R_out[h * 9 + k] = best < 0 ? NAN : P[best].R[k];
        for (int k = 0; k < 3; k++) t_out[h * 3 + k] = best < 0 ? NAN : P[best].t[k];
    }
    return 0;
}

}  // extern "C"
