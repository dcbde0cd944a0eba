// 6x6 FP64 helpers for the robust Gaussian fit (host side).
//
// Behavioural source: reference gpu-kernels/aux_funs.cpp:101-141, which uses cv::Matx66d::inv()/determinant
// (LU with partial pivoting) and trace().  OpenCV-free restatement; for the only size the pipeline uses
// (dims == 6, voldor/geometry.cpp:221) the results agree with OpenCV's LU to FP64 rounding and are narrowed to
// float by the caller.  For dims < 6 the matrix is treated as a proper dims x dims matrix (the reference
// reinterprets the buffer as 6x6 there, which is not a meaningful contract to keep).
#pragma once
#include <cmath>

namespace vb {
namespace linalg {

// LU with partial pivoting on an n x n copy (n <= 6); returns det, optionally the inverse.
inline double lu_det_inverse(const double* a_in, double* inv_out, int n) {
    double a[6][6], b[6][6];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
            a[i][j] = a_in[i * n + j];
            b[i][j] = (i == j) ? 1.0 : 0.0;
        }
    double det = 1.0;
    for (int i = 0; i < n; i++) {
        int k = i;
        for (int j = i + 1; j < n; j++)
            if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
        if (std::fabs(a[k][i]) < 2.220446049250313e-16 * 100) return 0.0;
        if (k != i) {
            for (int j = i; j < n; j++) {
                const double t = a[i][j];
                a[i][j] = a[k][j];
                a[k][j] = t;
            }
            for (int j = 0; j < n; j++) {
                const double t = b[i][j];
                b[i][j] = b[k][j];
                b[k][j] = t;
            }
            det = -det;
        }
        const double d = -1.0 / a[i][i];
        for (int j = i + 1; j < n; j++) {
            const double alpha = a[j][i] * d;
            for (int c = i + 1; c < n; c++) a[j][c] += alpha * a[i][c];
            for (int c = 0; c < n; c++) b[j][c] += alpha * b[i][c];
        }
        det *= a[i][i];
    }
    if (inv_out) {
        for (int i = n - 1; i >= 0; i--)
            for (int c = 0; c < n; c++) {
                double s = b[i][c];
                for (int k = i + 1; k < n; k++) s -= a[i][k] * b[k][c];
                b[i][c] = s / a[i][i];
            }
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) inv_out[i * n + j] = b[i][j];
    }
    return det;
}

// returns det; writes the inverse only when det > 0 (aux_funs.cpp:101-112)
inline double inverse6(const double* mat, double* mat_inv, int n) {
    const double det = lu_det_inverse(mat, nullptr, n);
    if (det > 0) lu_det_inverse(mat, mat_inv, n);
    return det;
}

// Ledoit-Wolf style shrinkage with a given intensity: S <- lambda*tr(S)/n*I + (1-lambda)*S (aux_funs.cpp:121-141)
inline void shrink_to_scaled_identity6(double* mat, double lambda, int n) {
    double tr = 0;
    for (int i = 0; i < n; i++) tr += mat[i * n + i];
    const double m = tr / (double)n;
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) mat[i * n + j] = lambda * m * (i == j ? 1.0 : 0.0) + (1 - lambda) * mat[i * n + j];
}

}  // namespace linalg
}  // namespace vb
