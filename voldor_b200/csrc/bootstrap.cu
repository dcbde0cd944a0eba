// Monocular bootstrap of a window that has no depth prior: relative pose of the first frame pair from the dense
// flow (essential matrix, least-median-of-squares) and a closed-form depth map.  Host code, runs once per sequence.
//
// Behavioural source: reference voldor/voldor.cpp:151-162 (bootstrap), voldor/geometry.cpp:288-332
// (estimate_camera_pose_epipolar: correspondences on a 4-pixel grid, cv::findEssentialMat(LMEDS, 0.999, 1.0),
// cv::recoverPose, then t := R*t) and :267-285 (estimate_depth_closed_form, clamp to [1e-2, 1e10]).
// The reference delegates the estimation to OpenCV (third-party arithmetic with its own RNG), so
// this is a functional replacement, validated by EM convergence, not a bit-parity component: normalised 8-point
// models inside an LMedS loop (deterministic xorshift sampling), rank-2 projection with equal singular values,
// inlier refit, and the cheirality test over the four (R, t) decompositions.
#include "bootstrap.h"
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

namespace vb {
namespace boot {
namespace {

// cyclic Jacobi eigen-decomposition of a symmetric n x n matrix (n <= 9); eigenvectors in the columns of V
void jacobi_eigen(int n, double* A, double* V, double* ev) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) V[i * n + j] = (i == j);
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) off += A[p * n + q] * A[p * n + q];
        if (off < 1e-30) break;
        for (int p = 0; p < n; p++)
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * n + q];
                if (std::fabs(apq) < 1e-300) continue;
                const double tau = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1 + tau * tau));
                const double c = 1 / std::sqrt(1 + t * t), s = t * c;
                for (int k = 0; k < n; k++) {
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq, A[k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; k++) {
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk, A[q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; k++) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq, V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < n; i++) ev[i] = A[i * n + i];
}

struct SVD3 {
    double U[9], S[3], V[9];
};
// E = U diag(S) V^T with S sorted descending, det(U) = det(V) = +1
SVD3 svd3(const double* E) {
    SVD3 r;
    double A[9], V[9], ev[3];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) A[i * 3 + j] = E[i] * E[j] + E[3 + i] * E[3 + j] + E[6 + i] * E[6 + j];
    jacobi_eigen(3, A, V, ev);
    int idx[3] = {0, 1, 2};
    std::sort(idx, idx + 3, [&](int a, int b) { return ev[a] > ev[b]; });
    for (int k = 0; k < 3; k++) {
        r.S[k] = std::sqrt(std::max(ev[idx[k]], 0.0));
        for (int i = 0; i < 3; i++) r.V[i * 3 + k] = V[i * 3 + idx[k]];
    }
    auto col = [&](double* M, int k, double* o) { o[0] = M[k], o[1] = M[3 + k], o[2] = M[6 + k]; };
    auto cross = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1], o[1] = a[2] * b[0] - a[0] * b[2], o[2] = a[0] * b[1] - a[1] * b[0];
    };
    double v0[3], v1[3], v2[3];
    col(r.V, 0, v0), col(r.V, 1, v1);
    cross(v0, v1, v2);
    for (int i = 0; i < 3; i++) r.V[i * 3 + 2] = v2[i];
    double u[3][3];
    for (int k = 0; k < 2; k++) {
        double vk[3];
        col(r.V, k, vk);
        for (int i = 0; i < 3; i++) u[k][i] = E[i * 3] * vk[0] + E[i * 3 + 1] * vk[1] + E[i * 3 + 2] * vk[2];
        const double n = std::sqrt(u[k][0] * u[k][0] + u[k][1] * u[k][1] + u[k][2] * u[k][2]);
        for (int i = 0; i < 3; i++) u[k][i] = n > 0 ? u[k][i] / n : (i == k);
    }
    cross(u[0], u[1], u[2]);
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) r.U[i * 3 + k] = u[k][i];
    return r;
}

struct Pt {
    double x1, y1, x2, y2;  // K-normalised coordinates
};

// 8-point (or more) essential matrix with Hartley conditioning; returns false on degenerate input
bool fit_essential(const std::vector<Pt>& pts, const std::vector<int>& sel, double* E) {
    const int n = (int)sel.size();
    double m1x = 0, m1y = 0, m2x = 0, m2y = 0;
    for (int i : sel) m1x += pts[i].x1, m1y += pts[i].y1, m2x += pts[i].x2, m2y += pts[i].y2;
    m1x /= n, m1y /= n, m2x /= n, m2y /= n;
    double d1 = 0, d2 = 0;
    for (int i : sel) {
        d1 += std::hypot(pts[i].x1 - m1x, pts[i].y1 - m1y);
        d2 += std::hypot(pts[i].x2 - m2x, pts[i].y2 - m2y);
    }
    if (d1 < 1e-12 || d2 < 1e-12) return false;
    const double s1 = std::sqrt(2.0) * n / d1, s2 = std::sqrt(2.0) * n / d2;
    double AtA[81] = {0};
    for (int i : sel) {
        const double a = (pts[i].x1 - m1x) * s1, b = (pts[i].y1 - m1y) * s1;
        const double c = (pts[i].x2 - m2x) * s2, d = (pts[i].y2 - m2y) * s2;
        const double r[9] = {c * a, c * b, c, d * a, d * b, d, a, b, 1};
        for (int p = 0; p < 9; p++)
            for (int q = 0; q < 9; q++) AtA[p * 9 + q] += r[p] * r[q];
    }
    double V[81], ev[9];
    jacobi_eigen(9, AtA, V, ev);
    int k = 0;
    for (int i = 1; i < 9; i++)
        if (ev[i] < ev[k]) k = i;
    double F[9];
    for (int i = 0; i < 9; i++) F[i] = V[i * 9 + k];
    // undo conditioning: E = T2^T F T1
    const double T1[9] = {s1, 0, -s1 * m1x, 0, s1, -s1 * m1y, 0, 0, 1}, T2[9] = {s2, 0, -s2 * m2x, 0, s2, -s2 * m2y, 0, 0, 1};
    double tmp[9], G[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) tmp[i * 3 + j] = F[i * 3] * T1[j] + F[i * 3 + 1] * T1[3 + j] + F[i * 3 + 2] * T1[6 + j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) G[i * 3 + j] = T2[i] * tmp[j] + T2[3 + i] * tmp[3 + j] + T2[6 + i] * tmp[6 + j];
    // project onto the essential manifold: singular values (1, 1, 0)
    const SVD3 s = svd3(G);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) E[i * 3 + j] = s.U[i * 3] * s.V[j * 3] + s.U[i * 3 + 1] * s.V[j * 3 + 1];
    return true;
}

inline double sampson(const double* E, const Pt& p) {
    const double Ex0 = E[0] * p.x1 + E[1] * p.y1 + E[2], Ex1 = E[3] * p.x1 + E[4] * p.y1 + E[5],
                 Ex2 = E[6] * p.x1 + E[7] * p.y1 + E[8];
    const double Et0 = E[0] * p.x2 + E[3] * p.y2 + E[6], Et1 = E[1] * p.x2 + E[4] * p.y2 + E[7];
    const double r = p.x2 * Ex0 + p.y2 * Ex1 + Ex2;
    return r * r / (Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1 + 1e-300);
}

}  // namespace

bool bootstrap_from_flow(const float* flow, int w, int h, const float* K9, float* R9, float* t3, float* depth) {
    const double fx = K9[0], cx = K9[2], fy = K9[4], cy = K9[5];
    const int step = 4;  // reference: geometry.h:17 sampling_2d_step = 4
    std::vector<Pt> pts;
    for (int y = 0; y < h; y += step)
        for (int x = 0; x < w; x += step) {
            const float* f = flow + ((size_t)y * w + x) * 2;
            if (!std::isfinite(f[0]) || !std::isfinite(f[1])) continue;
            pts.push_back({(x - cx) / fx, (y - cy) / fy, (x + f[0] - cx) / fx, (y + f[1] - cy) / fy});
        }
    const int n = (int)pts.size();
    if (n < 16) return false;

    // LMedS: confidence 0.999, assumed outlier ratio 0.45, 8-point samples
    const int iters = (int)std::ceil(std::log(1 - 0.999) / std::log(1 - std::pow(1 - 0.45, 8)));
    const int stride = std::max(1, n / 4000);  // scoring subset
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() {
        rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
        return rng;
    };
    double bestE[9] = {0}, best_med = 1e300;
    std::vector<double> err;
    std::vector<int> sel(8);
    for (int it = 0; it < iters; it++) {
        for (int k = 0; k < 8; k++) sel[k] = (int)(next() % n);
        double E[9];
        if (!fit_essential(pts, sel, E)) continue;
        err.clear();
        for (int i = 0; i < n; i += stride) err.push_back(sampson(E, pts[i]));
        std::nth_element(err.begin(), err.begin() + err.size() / 2, err.end());
        const double med = err[err.size() / 2];
        if (med < best_med) {
            best_med = med;
            for (int k = 0; k < 9; k++) bestE[k] = E[k];
        }
    }
    if (!(best_med < 1e300)) return false;
    // inlier refit (standard LMedS scale estimate)
    const double sigma = 2.5 * 1.4826 * (1 + 5.0 / (n - 8)) * std::sqrt(best_med);
    std::vector<int> inl;
    for (int i = 0; i < n; i++)
        if (sampson(bestE, pts[i]) <= sigma * sigma) inl.push_back(i);
    double E[9];
    if (inl.size() < 16 || !fit_essential(pts, inl, E))
        for (int k = 0; k < 9; k++) E[k] = bestE[k];

    // decomposition + cheirality (recoverPose)
    const SVD3 s = svd3(E);
    const double Wm[9] = {0, -1, 0, 1, 0, 0, 0, 0, 1};
    double Ra[9], Rb[9], UW[9], UWt[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            UW[i * 3 + j] = s.U[i * 3] * Wm[j] + s.U[i * 3 + 1] * Wm[3 + j] + s.U[i * 3 + 2] * Wm[6 + j];
            UWt[i * 3 + j] = s.U[i * 3] * Wm[j * 3] + s.U[i * 3 + 1] * Wm[j * 3 + 1] + s.U[i * 3 + 2] * Wm[j * 3 + 2];
        }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) {
            Ra[i * 3 + j] = UW[i * 3] * s.V[j * 3] + UW[i * 3 + 1] * s.V[j * 3 + 1] + UW[i * 3 + 2] * s.V[j * 3 + 2];
            Rb[i * 3 + j] = UWt[i * 3] * s.V[j * 3] + UWt[i * 3 + 1] * s.V[j * 3 + 1] + UWt[i * 3 + 2] * s.V[j * 3 + 2];
        }
    const double u3[3] = {s.U[2], s.U[5], s.U[8]};
    const double* Rc[2] = {Ra, Rb};
    int best = -1, best_cnt = -1;
    for (int c = 0; c < 4; c++) {
        const double* R = Rc[c >> 1];
        const double sg = (c & 1) ? -1.0 : 1.0;
        const double t[3] = {sg * u3[0], sg * u3[1], sg * u3[2]};
        int cnt = 0;
        for (int i : inl.empty() ? std::vector<int>() : inl) {
            const Pt& p = pts[i];
            // triangulate along ray 1: z1 * (R x1) + t = z2 * x2  ->  least squares for z1 from the cross product
            const double rx = R[0] * p.x1 + R[1] * p.y1 + R[2], ry = R[3] * p.x1 + R[4] * p.y1 + R[5],
                         rz = R[6] * p.x1 + R[7] * p.y1 + R[8];
            // (rx - x2 rz) z1 = x2 tz - tx ; (ry - y2 rz) z1 = y2 tz - ty
            const double a0 = rx - p.x2 * rz, a1 = ry - p.y2 * rz, b0 = p.x2 * t[2] - t[0], b1 = p.y2 * t[2] - t[1];
            const double z1 = (a0 * b0 + a1 * b1) / (a0 * a0 + a1 * a1 + 1e-300);
            const double z2 = z1 * rz + t[2];
            if (z1 > 0 && z2 > 0 && z1 < 50 && z2 < 50) cnt++;
        }
        if (cnt > best_cnt) best_cnt = cnt, best = c;
    }
    if (best < 0) return false;
    const double* R = Rc[best >> 1];
    const double sg = (best & 1) ? -1.0 : 1.0;
    float Rf[9], tf[3] = {(float)(sg * u3[0]), (float)(sg * u3[1]), (float)(sg * u3[2])};
    for (int i = 0; i < 9; i++) Rf[i] = (float)R[i];
    // reference: cam.t = cam.R * cam.t (geometry.cpp:330)
    for (int i = 0; i < 3; i++) t3[i] = Rf[i * 3] * tf[0] + Rf[i * 3 + 1] * tf[1] + Rf[i * 3 + 2] * tf[2];
    for (int i = 0; i < 9; i++) R9[i] = Rf[i];

    // closed-form depth (geometry.cpp:267-285)
    const float K[9] = {K9[0], K9[1], K9[2], K9[3], K9[4], K9[5], K9[6], K9[7], K9[8]};
    const float Kinv[9] = {1.f / K[0], 0, -K[2] / K[0], 0, 1.f / K[4], -K[5] / K[4], 0, 0, 1};
    float KR[9], KRK[9];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) KR[i * 3 + j] = K[i * 3] * Rf[j] + K[i * 3 + 1] * Rf[3 + j] + K[i * 3 + 2] * Rf[6 + j];
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            KRK[i * 3 + j] = KR[i * 3] * Kinv[j] + KR[i * 3 + 1] * Kinv[3 + j] + KR[i * 3 + 2] * Kinv[6 + j];
    const float b1 = K[0] * t3[0] + K[1] * t3[1] + K[2] * t3[2], b2 = K[3] * t3[0] + K[4] * t3[1] + K[5] * t3[2],
                b3 = K[6] * t3[0] + K[7] * t3[1] + K[8] * t3[2];
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            const float* f = flow + ((size_t)y * w + x) * 2;
            const float w1 = KRK[0] * x + KRK[1] * y + KRK[2], w2 = KRK[3] * x + KRK[4] * y + KRK[5],
                        w3 = KRK[6] * x + KRK[7] * y + KRK[8];
            const float a1 = x + f[0], a2 = y + f[1];
            const float nume = (a1 * b3 - b1) * (w1 - a1 * w3) + (a2 * b3 - b2) * (w2 - a2 * w3);
            const float deno = (w1 - a1 * w3) * (w1 - a1 * w3) + (w2 - a2 * w3) * (w2 - a2 * w3);
            depth[(size_t)y * w + x] = fminf(fmaxf(nume / deno, 1e-2f), 1e10f);
        }
    return true;
}

}  // namespace boot
}  // namespace vb
