// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// OpenCV-free stand-in for the three host helpers the reference's robust-Gaussian fit needs
// (reference: gpu-kernels/aux_funs.cpp:101-141, declared in gpu-kernels/aux_funs.h:3-7).  The reference
// implements them with cv::Matx66d (determinant / inv() = LU with partial pivoting / trace); OpenCV C++
// headers are not installed in this image, so the reference .cu files are linked against this shim when
// the oracle library oracle/_ref/libgpu_kernels_ref.so is built.  This is the single deviation of the
// oracle from the reference sources (agreement with OpenCV's LU is ~1e-15 relative, then narrowed to float).
//
// Contract kept from the reference:
//   * the matrix is always treated as 6x6 regardless of `dims` (aux_funs.cpp:102,113,122 use Matx66d);
//   * inverse() returns the determinant and writes the inverse only when det > 0 (aux_funs.cpp:104-111);
//   * regularize: S* = lambda * tr(S)/dims * I + (1-lambda) * S, returns det(S*) (aux_funs.cpp:121-141).
#include <cmath>
#include <cstring>

namespace {
const int M = 6;

// LU decomposition with partial pivoting on a 6x6 copy; returns det, optionally solves for the inverse.
double lu6(const double* a_in, double* inv_out) {
    double a[M][M];
    double b[M][M];
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) {
            a[i][j] = a_in[i * M + j];
            b[i][j] = (i == j) ? 1.0 : 0.0;
        }
    double det = 1.0;
    for (int i = 0; i < M; i++) {
        int k = i;
        for (int j = i + 1; j < M; j++)
            if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
        if (std::fabs(a[k][i]) < 2.220446049250313e-16 * 100) return 0.0;
        if (k != i) {
            for (int j = i; j < M; j++) { double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
            for (int j = 0; j < M; j++) { double t = b[i][j]; b[i][j] = b[k][j]; b[k][j] = t; }
            det = -det;
        }
        double d = -1.0 / a[i][i];
        for (int j = i + 1; j < M; j++) {
            double alpha = a[j][i] * d;
            for (int c = i + 1; c < M; c++) a[j][c] += alpha * a[i][c];
            for (int c = 0; c < M; c++) b[j][c] += alpha * b[i][c];
        }
        det *= a[i][i];
    }
    if (inv_out) {
        // back substitution
        for (int i = M - 1; i >= 0; i--) {
            for (int c = 0; c < M; c++) {
                double s = b[i][c];
                for (int k = i + 1; k < M; k++) s -= a[i][k] * b[k][c];
                b[i][c] = s / a[i][i];
            }
        }
        for (int i = 0; i < M; i++)
            for (int j = 0; j < M; j++) inv_out[i * M + j] = b[i][j];
    }
    return det;
}
}  // namespace

double determinant(double* mat, int /*N*/) { return lu6(mat, nullptr); }

double inverse(double* mat, double* mat_inv, int N) {
    double det = lu6(mat, nullptr);
    if (det > 0) {
        double inv[M * M];
        lu6(mat, inv);
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) mat_inv[i * N + j] = inv[i * M + j];
    }
    return det;
}

double regularize_covar_LW_given_lambda(double* mat, double* mat_ret, double lambda, int dims) {
    double tr = 0;
    for (int i = 0; i < M; i++) tr += mat[i * M + i];
    double m = tr / (double)dims;
    double s[M * M];
    for (int i = 0; i < M; i++)
        for (int j = 0; j < M; j++) s[i * M + j] = lambda * m * (i == j ? 1.0 : 0.0) + (1 - lambda) * mat[i * M + j];
    for (int i = 0; i < dims; i++)
        for (int j = 0; j < dims; j++) mat_ret[i * dims + j] = s[i * M + j];
    return lu6(s, nullptr);
}
