// Monocular bootstrap of a window that has no depth prior: relative pose of the first frame pair from the dense
// flow (essential matrix, least-median-of-squares) and a closed-form depth map.  Host code, runs once per sequence.
//
// Behavioural source: reference voldor/voldor.cpp:151-162 (bootstrap), voldor/geometry.cpp:288-332
// (estimate_camera_pose_epipolar: correspondences on a 4-pixel grid, cv::findEssentialMat(LMEDS, 0.999, 1.0),
// cv::recoverPose, then t := R*t) and :267-285 (estimate_depth_closed_form, clamp to [1e-2, 1e10]).
// The reference delegates the estimation to OpenCV (third-party arithmetic with its own RNG), so
// this is a functional replacement, validated against synthetic ground truth and by EM convergence, not a
// bit-parity component: an LMedS loop (deterministic xorshift sampling) over calibrated 5-dof (R, t) models fitted to
// 8-point samples by Levenberg-Marquardt on the Sampson error, inlier refit, cheirality test for the sign of t.
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

struct Pt {
    double x1, y1, x2, y2;  // K-normalised coordinates
};

// calibrated two-view model: x2 ~ R x1 + t, |t| = 1
struct Model {
    double R[9], t[3];
};

inline void essential(const Model& m, double* E) {
    const double* t = m.t;
    const double Tx[9] = {0, -t[2], t[1], t[2], 0, -t[0], -t[1], t[0], 0};
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) E[i * 3 + j] = Tx[i * 3] * m.R[j] + Tx[i * 3 + 1] * m.R[3 + j] + Tx[i * 3 + 2] * m.R[6 + j];
}

// signed Sampson residual of one correspondence
inline double sampson(const double* E, const Pt& p) {
    const double Ex0 = E[0] * p.x1 + E[1] * p.y1 + E[2], Ex1 = E[3] * p.x1 + E[4] * p.y1 + E[5],
                 Ex2 = E[6] * p.x1 + E[7] * p.y1 + E[8];
    const double Et0 = E[0] * p.x2 + E[3] * p.y2 + E[6], Et1 = E[1] * p.x2 + E[4] * p.y2 + E[7];
    const double r = p.x2 * Ex0 + p.y2 * Ex1 + Ex2;
    return r / std::sqrt(Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1 + 1e-300);
}

// translation direction for a fixed rotation: t . (R x1 x x2) = 0 for every correspondence
void linear_translation(Model& m, const std::vector<Pt>& pts, const std::vector<int>& sel) {
    double A[9] = {0}, V[9], ev[3];
    for (int i : sel) {
        const Pt& p = pts[i];
        const double r[3] = {m.R[0] * p.x1 + m.R[1] * p.y1 + m.R[2], m.R[3] * p.x1 + m.R[4] * p.y1 + m.R[5],
                             m.R[6] * p.x1 + m.R[7] * p.y1 + m.R[8]};
        const double c[3] = {r[1] - r[2] * p.y2, r[2] * p.x2 - r[0], r[0] * p.y2 - r[1] * p.x2};
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) A[a * 3 + b] += c[a] * c[b];
    }
    jacobi_eigen(3, A, V, ev);
    int k = 0;
    for (int i = 1; i < 3; i++)
        if (ev[i] < ev[k]) k = i;
    for (int i = 0; i < 3; i++) m.t[i] = V[i * 3 + k];
}

// local 5-dof update: rotation by exp(d[0..2]) on the left, translation moved in its tangent plane by d[3..4]
Model retract(const Model& m, const double* d) {
    Model o;
    const double th = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    double dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (th > 1e-14) {
        const double k[3] = {d[0] / th, d[1] / th, d[2] / th}, s = std::sin(th), c = 1 - std::cos(th);
        const double Kx[9] = {0, -k[2], k[1], k[2], 0, -k[0], -k[1], k[0], 0};
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++)
                dR[i * 3 + j] += s * Kx[i * 3 + j] + c * (Kx[i * 3] * Kx[j] + Kx[i * 3 + 1] * Kx[3 + j] + Kx[i * 3 + 2] * Kx[6 + j]);
    }
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) o.R[i * 3 + j] = dR[i * 3] * m.R[j] + dR[i * 3 + 1] * m.R[3 + j] + dR[i * 3 + 2] * m.R[6 + j];
    // tangent basis of the unit sphere at t
    const double* t = m.t;
    int a = std::fabs(t[0]) < std::fabs(t[1]) ? (std::fabs(t[0]) < std::fabs(t[2]) ? 0 : 2) : (std::fabs(t[1]) < std::fabs(t[2]) ? 1 : 2);
    double e[3] = {0, 0, 0}, b1[3], b2[3];
    e[a] = 1;
    b1[0] = t[1] * e[2] - t[2] * e[1], b1[1] = t[2] * e[0] - t[0] * e[2], b1[2] = t[0] * e[1] - t[1] * e[0];
    const double n1 = std::sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
    for (int i = 0; i < 3; i++) b1[i] /= n1;
    b2[0] = t[1] * b1[2] - t[2] * b1[1], b2[1] = t[2] * b1[0] - t[0] * b1[2], b2[2] = t[0] * b1[1] - t[1] * b1[0];
    double nt = 0;
    for (int i = 0; i < 3; i++) o.t[i] = t[i] + d[3] * b1[i] + d[4] * b2[i], nt += o.t[i] * o.t[i];
    nt = std::sqrt(nt);
    for (int i = 0; i < 3; i++) o.t[i] /= nt;
    return o;
}

double sum_sq(const Model& m, const std::vector<Pt>& pts, const std::vector<int>& sel, std::vector<double>* r = nullptr) {
    double E[9], acc = 0;
    essential(m, E);
    if (r) r->resize(sel.size());
    for (size_t i = 0; i < sel.size(); i++) {
        const double v = sampson(E, pts[sel[i]]);
        acc += v * v;
        if (r) (*r)[i] = v;
    }
    return acc;
}

// Levenberg-Marquardt on the Sampson residuals of the selected correspondences (forward-difference Jacobian)
void refine(Model& m, const std::vector<Pt>& pts, const std::vector<int>& sel, int iters) {
    std::vector<double> r0, r1, J(sel.size() * 5);
    double lambda = 1e-4, cost = sum_sq(m, pts, sel, &r0);
    for (int it = 0; it < iters; it++) {
        const double h = 1e-6;
        for (int k = 0; k < 5; k++) {
            double d[5] = {0, 0, 0, 0, 0};
            d[k] = h;
            sum_sq(retract(m, d), pts, sel, &r1);
            for (size_t i = 0; i < sel.size(); i++) J[i * 5 + k] = (r1[i] - r0[i]) / h;
        }
        double H[25] = {0}, g[5] = {0};
        for (size_t i = 0; i < sel.size(); i++)
            for (int a = 0; a < 5; a++) {
                g[a] += J[i * 5 + a] * r0[i];
                for (int b = 0; b < 5; b++) H[a * 5 + b] += J[i * 5 + a] * J[i * 5 + b];
            }
        bool improved = false;
        for (int attempt = 0; attempt < 6 && !improved; attempt++) {
            double A[5][6];
            for (int a = 0; a < 5; a++) {
                for (int b = 0; b < 5; b++) A[a][b] = H[a * 5 + b] + (a == b ? lambda * (H[a * 5 + a] + 1e-12) : 0);
                A[a][5] = -g[a];
            }
            bool ok = true;
            for (int c = 0; c < 5 && ok; c++) {
                int piv = c;
                for (int rr = c + 1; rr < 5; rr++)
                    if (std::fabs(A[rr][c]) > std::fabs(A[piv][c])) piv = rr;
                if (std::fabs(A[piv][c]) < 1e-300) ok = false;
                for (int k = 0; k < 6; k++) std::swap(A[c][k], A[piv][k]);
                for (int rr = 0; rr < 5 && ok; rr++) {
                    if (rr == c) continue;
                    const double f = A[rr][c] / A[c][c];
                    for (int k = c; k < 6; k++) A[rr][k] -= f * A[c][k];
                }
            }
            if (!ok) {
                lambda *= 10;
                continue;
            }
            double d[5];
            for (int a = 0; a < 5; a++) d[a] = A[a][5] / A[a][a];
            const Model cand = retract(m, d);
            const double c2 = sum_sq(cand, pts, sel, &r1);
            if (c2 < cost) {
                m = cand, cost = c2, r0.swap(r1), lambda = std::max(lambda * 0.3, 1e-9), improved = true;
            } else {
                lambda *= 10;
            }
        }
        if (!improved) break;
    }
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

    // LMedS over 8-point samples of the calibrated 5-dof model.  Every sample starts from R = I (successive video
    // frames), takes the translation from the linear constraint and polishes (R, t) on the sample; unlike the
    // linear 8-point fit this stays well posed on near-planar scenes.  0.999 confidence at 45 % outliers.
    const int iters = (int)std::ceil(std::log(1 - 0.999) / std::log(1 - std::pow(1 - 0.45, 8)));
    const int stride = std::max(1, n / 4000);  // scoring subset
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() {
        rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
        return rng;
    };
    Model best_m{};
    double best_med = 1e300;
    std::vector<double> err;
    std::vector<int> sel(8);
    for (int it = 0; it < iters; it++) {
        for (int k = 0; k < 8; k++) sel[k] = (int)(next() % n);
        Model m{{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 0, 1}};
        linear_translation(m, pts, sel);
        refine(m, pts, sel, 6);
        double E[9];
        essential(m, E);
        err.clear();
        for (int i = 0; i < n; i += stride) {
            const double r = sampson(E, pts[i]);
            err.push_back(r * r);
        }
        std::nth_element(err.begin(), err.begin() + err.size() / 2, err.end());
        const double med = err[err.size() / 2];
        if (med < best_med) best_med = med, best_m = m;
    }
    if (!(best_med < 1e300)) return false;
    // inliers by the standard LMedS scale estimate, then a fit on all of them
    const double sigma = 2.5 * 1.4826 * (1 + 5.0 / (n - 8)) * std::sqrt(best_med);
    std::vector<int> inl;
    {
        double E[9];
        essential(best_m, E);
        for (int i = 0; i < n; i++)
            if (std::fabs(sampson(E, pts[i])) <= sigma) inl.push_back(i);
    }
    if (inl.size() >= 16) refine(best_m, pts, inl, 15);

    // sign of t by cheirality (the rotation is the one next to identity; its twisted pair turns by pi)
    const double* R = best_m.R;
    int cnt[2] = {0, 0};
    for (int sgi = 0; sgi < 2; sgi++) {
        const double sg = sgi ? -1.0 : 1.0;
        const double t[3] = {sg * best_m.t[0], sg * best_m.t[1], sg * best_m.t[2]};
        for (int i : inl) {
            const Pt& p = pts[i];
            const double rx = R[0] * p.x1 + R[1] * p.y1 + R[2], ry = R[3] * p.x1 + R[4] * p.y1 + R[5],
                         rz = R[6] * p.x1 + R[7] * p.y1 + R[8];
            // z1 (R x1) + t = z2 x2  ->  (rx - x2 rz) z1 = x2 tz - tx ; (ry - y2 rz) z1 = y2 tz - ty
            const double a0 = rx - p.x2 * rz, a1 = ry - p.y2 * rz, b0 = p.x2 * t[2] - t[0], b1 = p.y2 * t[2] - t[1];
            const double z1 = (a0 * b0 + a1 * b1) / (a0 * a0 + a1 * a1 + 1e-300);
            const double z2 = z1 * rz + t[2];
            if (z1 > 0 && z2 > 0) cnt[sgi]++;
        }
    }
    const double sg = cnt[1] > cnt[0] ? -1.0 : 1.0;
    const double u3[3] = {best_m.t[0], best_m.t[1], best_m.t[2]};
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
