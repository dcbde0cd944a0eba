// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// Out-of-line part of the OpenCV stand-in (cv_min/opencv2/core_min.hpp) that lets the reference's host sources compile
// unmodified.  Compiled with -ffp-contract=off: every expression below rounds where it is written.
#include "cv_min/opencv2/core_min.hpp"
#include "../../voldor_b200/csrc/host_math.h"

namespace cv {

[[noreturn]] void cvmin_unsupported(const char* what) {
    fprintf(stderr, "cv_min: '%s' is not available in the OpenCV stand-in (oracle/ref_shim/cv_min)\n", what);
    abort();
}

namespace {
// monocular bootstrap handed in by the test: what findEssentialMat + recoverPose would have produced
double g_epi_R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
double g_epi_t[3] = {0, 0, 0};
bool g_epi_set = false;

template <typename F>
void for_each_f32(const Mat& m, F f) {
    if (m.type() != CV_32F) cvmin_unsupported("element-wise operation on a non-CV_32F matrix");
    for (int y = 0; y < m.rows; y++) {
        float* row = (float*)(m.data + m.step * y);
        for (int x = 0; x < m.cols; x++) f(row[x]);
    }
}
}  // namespace

extern "C" void cvmin_inject_epipolar(const float* R9, const float* t3) {
    g_epi_set = R9 != nullptr;
    if (!R9) return;
    for (int i = 0; i < 9; i++) g_epi_R[i] = R9[i];
    for (int i = 0; i < 3; i++) g_epi_t[i] = t3[i];
}

Mat Mat::ones(int r, int c, int type) {
    Mat m(r, c, type);
    for_each_f32(m, [](float& v) { v = 1.f; });
    return m;
}

Mat Mat::eye(int r, int c, int type) {
    Mat m = zeros(r, c, type);
    for (int i = 0; i < std::min(r, c); i++) {
        if (type == CV_32F)
            m.at<float>(i, i) = 1.f;
        else if (type == CV_64F)
            m.at<double>(i, i) = 1.;
        else
            cvmin_unsupported("Mat::eye of this type");
    }
    return m;
}

Mat& Mat::operator=(const Scalar& s) {
    const float v = (float)s[0];
    for_each_f32(*this, [v](float& e) { e = v; });
    return *this;
}

void Mat::convertTo(Mat& dst, int rtype, double alpha, double beta) const {
    const int dt = rtype < 0 ? depth() : (rtype & 7);
    if (channels() != 1 || (depth() != CV_32F && depth() != CV_64F) || (dt != CV_32F && dt != CV_64F))
        cvmin_unsupported("Mat::convertTo outside single-channel 32F/64F");
    Mat out(rows, cols, dt);
    for (int y = 0; y < rows; y++)
        for (int x = 0; x < cols; x++) {
            const double v = depth() == CV_32F ? (double)at<float>(y, x) : at<double>(y, x);
            const double r = alpha == 1 && beta == 0 ? v : v * alpha + beta;
            if (dt == CV_32F)
                out.at<float>(y, x) = (float)r;
            else
                out.at<double>(y, x) = r;
        }
    dst = out;
}

Mat Mat::diag() const {
    const int n = std::min(rows, cols);
    Mat d(n, 1, type_);
    for (int i = 0; i < n; i++) memcpy(d.data + d.step * i, data + step * i + elemSize() * i, elemSize());
    return d;
}

// 3x3 only: cofactors and determinant in double, one rounding per element (what cv::invert does for small float matrices)
Mat Mat::inv() const {
    if (rows != 3 || cols != 3 || type_ != CV_32F) cvmin_unsupported("Mat::inv outside 3x3 CV_32F");
    double a[9];
    for (int i = 0; i < 9; i++) a[i] = at<float>(i / 3, i % 3);
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    Mat r = zeros(3, 3, CV_32F);
    if (det == 0) return r;
    const double d = 1. / det;
    const double adj[9] = {c00, a[2] * a[7] - a[1] * a[8], a[1] * a[5] - a[2] * a[4],
                           c01, a[0] * a[8] - a[2] * a[6], a[2] * a[3] - a[0] * a[5],
                           c02, a[1] * a[6] - a[0] * a[7], a[0] * a[4] - a[1] * a[3]};
    for (int i = 0; i < 9; i++) r.at<float>(i / 3, i % 3) = (float)(adj[i] * d);
    return r;
}

Mat operator*(const Mat& a, const Mat& b) {
    if (a.cols != b.rows || a.type() != b.type() || (a.type() != CV_32F && a.type() != CV_64F))
        cvmin_unsupported("matrix product of these shapes / types");
    Mat c(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; i++)
        for (int j = 0; j < b.cols; j++) {
            double s = 0;  // cv::gemm accumulates float products in double
            for (int k = 0; k < a.cols; k++)
                s += a.type() == CV_32F ? (double)a.at<float>(i, k) * (double)b.at<float>(k, j)
                                        : a.at<double>(i, k) * b.at<double>(k, j);
            if (a.type() == CV_32F)
                c.at<float>(i, j) = (float)s;
            else
                c.at<double>(i, j) = s;
        }
    return c;
}

Mat& operator*=(Mat& a, double s) {
    const float f = (float)s;
    for_each_f32(a, [f](float& v) { v = v * f; });
    return a;
}
Mat& operator/=(Mat& a, double s) {
    const float f = (float)(1. / s);
    for_each_f32(a, [f](float& v) { v = v * f; });
    return a;
}
Mat& operator+=(Mat& a, const Mat& b) {
    if (a.rows != b.rows || a.cols != b.cols || a.type() != CV_32F || b.type() != CV_32F)
        cvmin_unsupported("Mat += Mat of these shapes / types");
    for (int y = 0; y < a.rows; y++)
        for (int x = 0; x < a.cols; x++) a.at<float>(y, x) = a.at<float>(y, x) + b.at<float>(y, x);
    return a;
}
Mat operator*(const Mat& a, double s) {
    Mat r = a.clone();
    r *= s;
    return r;
}
Mat operator*(double s, const Mat& a) { return a * s; }
Mat operator/(const Mat& a, double s) {
    Mat r = a.clone();
    r /= s;
    return r;
}
Mat operator/(double s, const Mat& a) {
    Mat r = a.clone();
    const float f = (float)s;
    for_each_f32(r, [f](float& v) { v = v == 0.f ? 0.f : f / v; });
    return r;
}

Scalar sum(const Mat& m) {
    double s = 0;
    for_each_f32(m, [&s](float& v) { s += v; });
    return Scalar(s);
}
Scalar mean(const Mat& m) { return Scalar(m.total() ? sum(m)[0] / (double)m.total() : 0.); }
double norm(const Mat& m, int) {
    double s = 0;
    for_each_f32(m, [&s](float& v) { s += (double)v * v; });
    return std::sqrt(s);
}
bool checkRange(const Mat& m) {
    bool ok = true;
    for_each_f32(m, [&ok](float& v) { ok = ok && std::isfinite(v); });
    return ok;
}

void Rodrigues(const Mat& R, Vec3f& rvec) {
    if (R.rows != 3 || R.cols != 3 || R.type() != CV_32F) cvmin_unsupported("Rodrigues(matrix) outside 3x3 CV_32F");
    float r9[9];
    for (int i = 0; i < 9; i++) r9[i] = R.at<float>(i / 3, i % 3);
    vb::hm::matrix_to_rvec(r9, rvec.val);
}
void Rodrigues(const Matx33f& R, Vec3f& rvec) { vb::hm::matrix_to_rvec(R.val, rvec.val); }
void Rodrigues(const Vec3f& rvec, Mat& R) {
    float r9[9];
    vb::hm::rvec_to_matrix(rvec.val, r9);
    R.create(3, 3, CV_32F);
    for (int i = 0; i < 9; i++) R.at<float>(i / 3, i % 3) = r9[i];
}

Mat findEssentialMat(const Mat&, const Mat&, const Mat&, int, double, double, Mat&) {
    if (!g_epi_set) cvmin_unsupported("findEssentialMat without an injected pose (cvmin_inject_epipolar)");
    return Mat::zeros(3, 3, CV_64F);
}
int recoverPose(const Mat&, const Mat&, const Mat&, const Mat&, Mat& R, Mat& t) {
    if (!g_epi_set) cvmin_unsupported("recoverPose without an injected pose (cvmin_inject_epipolar)");
    R.create(3, 3, CV_64F);
    t.create(3, 1, CV_64F);
    for (int i = 0; i < 9; i++) R.at<double>(i / 3, i % 3) = g_epi_R[i];
    for (int i = 0; i < 3; i++) t.at<double>(i, 0) = g_epi_t[i];
    return 0;
}

bool eigen(const Matx33f&, Matx31f&, Matx33f&) { cvmin_unsupported("eigen (KITTI ground plane, deprecated in the reference)"); }
bool solvePnP(const _InputArray&, const _InputArray&, const Mat&, const Mat&, Vec3d&, Vec3d&, bool, int) {
    cvmin_unsupported("solvePnP (--cpu_p3p without --lambdatwist)");
}
void resize(const Mat&, Mat&, Size, double, double) { cvmin_unsupported("resize (--resize_factor != 1)"); }
void split(const Mat&, Mat*) { cvmin_unsupported("split"); }
void merge(const std::vector<Mat>&, Mat&) { cvmin_unsupported("merge"); }
void cartToPolar(const Mat&, const Mat&, Mat&, Mat&, bool) { cvmin_unsupported("cartToPolar"); }
void normalize(const Mat&, Mat&, double, double, int) { cvmin_unsupported("normalize"); }
void cvtColor(const Mat&, Mat&, int) { cvmin_unsupported("cvtColor"); }
void imshow(const std::string&, const Mat&) { cvmin_unsupported("imshow (--debug)"); }
bool imwrite(const std::string&, const Mat&) { cvmin_unsupported("imwrite"); }
int waitKey(int) { return -1; }
void destroyAllWindows() {}

}  // namespace cv

// ---- self-description of the stand-in's Mat semantics for tests/test_cpu_reference_host.py: the handful of OpenCV
// behaviours the reference's host code leans on (views write through, headers are shallow, clone is deep, at<T>(i, j)
// indexes in units of T, the comma initialiser fills row-major, K.inv(), small products, scalar assignment in place).
extern "C" int cvmin_selftest(double* out) {
    using namespace cv;
    int k = 0;
    Mat pool(4, 6, CV_32F);
    for (int i = 0; i < 24; i++) pool.at<float>(i / 6, i % 6) = (float)i;
    Mat alias = pool;                      // shallow
    Mat copy = pool.clone();               // deep
    pool.colRange(0, 3) *= 2.0;            // writes through the view (geometry.cpp:187)
    out[k++] = alias.at<float>(1, 2);      // 16: doubled through the alias
    out[k++] = alias.at<float>(1, 3);      // 9: outside the column range
    out[k++] = copy.at<float>(1, 2);       // 8: the clone kept the old value
    pool.at<Vec3f>(2, 1) = Vec3f(-1, -2, -3);  // second Vec3f of row 2 = columns 3..5 (geometry.cpp:160)
    out[k++] = pool.at<float>(2, 3), out[k++] = pool.at<float>(2, 5);
    Mat head = pool.rowRange(0, 2);        // geometry.cpp:176
    out[k++] = head.rows, out[k++] = head.at<float>(1, 5);
    Mat row(1, 6, CV_32F);
    row.at<Vec3f>(1) = Vec3f(7, 8, 9);     // single index on a row vector, in units of Vec3f (geometry.cpp:181)
    out[k++] = row.at<float>(0, 3), out[k++] = row.at<float>(0, 5);
    Mat t = Mat::zeros(3, 1, CV_32F);
    t.at<Vec3f>(0) = Vec3f(1, 2, 3);       // voldor.cpp:64
    out[k++] = t.at<float>(2);
    Mat K = (Mat_<float>(3, 3) << 500, 0, 320, 0, 510, 240, 0, 0, 1);
    out[k++] = K.at<float>(1, 2);          // 240: row-major fill
    Mat Ki = K.inv();
    Mat I = K * Ki;
    out[k++] = Ki.at<float>(0, 0), out[k++] = Ki.at<float>(0, 2), out[k++] = I.at<float>(0, 0) + I.at<float>(1, 1) + I.at<float>(2, 2);
    Mat c = Mat::ones(2, 2, CV_32F);
    Mat c2 = c;
    c = 0;                                 // in place: the second header sees it (geometry.cpp:204)
    out[k++] = c2.at<float>(1, 1);
    Mat d64 = Mat::eye(3, 3, CV_64F);
    d64.convertTo(d64, CV_32F);            // geometry.cpp:329
    out[k++] = d64.type() == CV_32F ? d64.at<float>(2, 2) : -1;
    Mat s = Mat::ones(3, 2, CV_32F);
    out[k++] = sum(s)[0], out[k++] = norm(s), out[k++] = checkRange(s) ? 1 : 0;
    s.at<float>(2, 1) = NAN;
    out[k++] = checkRange(s) ? 1 : 0;
    Mat q = 10.0 / Mat::zeros(1, 2, CV_32F);  // s / m is 0 where m is 0 (voldor.cpp:33 on missing disparities)
    out[k++] = q.at<float>(0, 1);
    return k;
}
