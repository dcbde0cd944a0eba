// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// A small, self-written stand-in for the part of OpenCV's C++ API that the reference's HOST code uses
// (voldor/utils.h, config.h, voldor.cpp, geometry.cpp, py_export.cpp, utils.cpp), so that those files compile
// UNMODIFIED, from where they lie, into oracle/_ref/libvoldor_host_ref.so (OpenCV's C++ library is not in this image).
// The reference's control flow — iteration loop, truncation logic, the NULL-cached call patterns of every kernel, host
// compaction and scaling between the calls — then runs as the reference's own compiled code.
//
// What is real here and what is not:
//   * Mat is a reference-counted header over a byte buffer with row/column views and shallow copies, like cv::Mat;
//     at<T>(), rowRange/colRange, clone, convertTo (32F <-> 64F), scalar and element-wise operators, small dense
//     products, a 3x3 inverse, sum / mean / norm / checkRange follow OpenCV's documented semantics.  Arithmetic that
//     reaches the compared outputs is pinned by the parity tests to the same formulas the product uses
//     (csrc/host_math.h): scaling by s multiplies by (float)s, division by s multiplies by (float)(1/s), Rodrigues is
//     host_math's (itself pinned to cv2 in tests/test_cpu_host_math.py).
//   * findEssentialMat / recoverPose return a pose injected by the test (cvmin_inject_epipolar): OpenCV's 5-point
//     LMedS solver is not restated (same treatment as everywhere in this repository).
//   * solvePnP (the --cpu_p3p path with AP3P), eigen (deprecated KITTI ground plane), resize (resize_factor != 1) and
//     the GUI / image-file calls are declared and abort with a message when reached.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <typeinfo>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

enum { NORM_L2 = 4, NORM_MINMAX = 32 };
enum { RANSAC = 8, LMEDS = 4 };
enum { COLOR_HSV2BGR = 54 };

[[noreturn]] void cvmin_unsupported(const char* what);

struct Scalar {
    double val[4];
    Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : val{a, b, c, d} {}
    double operator[](int i) const { return val[i]; }
};

struct Size {
    int width = 0, height = 0;
    Size() {}
    Size(int w, int h) : width(w), height(h) {}
};

struct Rect {
    int x = 0, y = 0, width = 0, height = 0;
    Rect() {}
    Rect(int x_, int y_, int w_, int h_) : x(x_), y(y_), width(w_), height(h_) {}
};

template <typename T, int n>
struct Vec {
    T val[n];
    Vec() {
        for (int i = 0; i < n; i++) val[i] = T(0);
    }
    Vec(T a, T b) : Vec() { val[0] = a, val[1] = b; }
    Vec(T a, T b, T c) : Vec() { val[0] = a, val[1] = b, val[2] = c; }
    Vec(T a, T b, T c, T d) : Vec() { val[0] = a, val[1] = b, val[2] = c, val[3] = d; }
    Vec(T a, T b, T c, T d, T e, T f) : Vec() { val[0] = a, val[1] = b, val[2] = c, val[3] = d, val[4] = e, val[5] = f; }
    explicit Vec(const T* p) {
        for (int i = 0; i < n; i++) val[i] = p[i];
    }
    template <typename U>
    operator Vec<U, n>() const {
        Vec<U, n> r;
        for (int i = 0; i < n; i++) r.val[i] = (U)val[i];
        return r;
    }
    T& operator[](int i) { return val[i]; }
    const T& operator[](int i) const { return val[i]; }
    T dot(const Vec& o) const {
        T s = T(0);
        for (int i = 0; i < n; i++) s += val[i] * o.val[i];
        return s;
    }
    Vec operator-() const {
        Vec r;
        for (int i = 0; i < n; i++) r.val[i] = -val[i];
        return r;
    }
    Vec& operator/=(double s) {
        for (int i = 0; i < n; i++) val[i] = (T)(val[i] / s);
        return *this;
    }
    Vec& operator*=(double s) {
        for (int i = 0; i < n; i++) val[i] = (T)(val[i] * s);
        return *this;
    }
};
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<float, 6> Vec6f;
typedef Vec<double, 3> Vec3d;

template <typename T, int n>
std::ostream& operator<<(std::ostream& os, const Vec<T, n>& v) {
    os << "[";
    for (int i = 0; i < n; i++) os << v.val[i] << (i + 1 < n ? ", " : "]");
    return os;
}

template <typename T>
struct Point_ {
    T x = T(0), y = T(0);
    Point_() {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
};
typedef Point_<float> Point2f;

template <typename T>
struct Point3_ {
    T x = T(0), y = T(0), z = T(0);
    Point3_() {}
    Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
    T dot(const Point3_& o) const { return x * o.x + y * o.y + z * o.z; }
    operator Vec<T, 3>() const { return Vec<T, 3>(x, y, z); }
    Point3_& operator+=(const Point3_& o) {
        x += o.x, y += o.y, z += o.z;
        return *this;
    }
    Point3_& operator/=(double s) {
        x = (T)(x / s), y = (T)(y / s), z = (T)(z / s);
        return *this;
    }
};
typedef Point3_<float> Point3f;
template <typename T>
Point3_<T> operator*(const Point3_<T>& p, float s) {
    return Point3_<T>(p.x * s, p.y * s, p.z * s);
}

template <typename T, int m, int n>
struct Matx {
    T val[m * n];
    Matx() {
        for (int i = 0; i < m * n; i++) val[i] = T(0);
    }
    explicit Matx(const T* p) {
        for (int i = 0; i < m * n; i++) val[i] = p[i];
    }
    Matx(T a0, T a1, T a2, T a3, T a4, T a5, T a6, T a7, T a8) {
        static_assert(m * n == 9, "nine-value constructor is for 3x3");
        const T v[9] = {a0, a1, a2, a3, a4, a5, a6, a7, a8};
        for (int i = 0; i < 9; i++) val[i] = v[i];
    }
};
typedef Matx<float, 3, 3> Matx33f;
typedef Matx<float, 3, 1> Matx31f;

// ---------------------------------------------------------------------------------------------- Mat
struct Mat {
    int type_ = CV_32F;
    int rows = 0, cols = 0;
    size_t step = 0;  // bytes between rows
    unsigned char* data = nullptr;
    std::shared_ptr<unsigned char> owner;  // null for headers over caller memory

    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(Size s, int type) { create(s.height, s.width, type); }
    Mat(int r, int c, int type, void* p) : type_(type), rows(r), cols(c), step((size_t)c * esz(type)), data((unsigned char*)p) {}
    Mat(Size s, int type, void* p) : Mat(s.height, s.width, type, p) {}

    static size_t esz1(int type) { return (type & 7) == CV_64F ? 8 : (type & 7) == CV_32F ? 4 : 1; }
    static int chans(int type) { return (type >> 3) + 1; }
    static size_t esz(int type) { return esz1(type) * chans(type); }

    void create(int r, int c, int type) {
        if (data && rows == r && cols == c && type_ == type && step == (size_t)c * esz(type)) return;
        type_ = type, rows = r, cols = c, step = (size_t)c * esz(type);
        const size_t bytes = std::max<size_t>((size_t)r * step, 1);
        owner.reset(new unsigned char[bytes], std::default_delete<unsigned char[]>());
        data = owner.get();
    }
    int type() const { return type_; }
    int depth() const { return type_ & 7; }
    int channels() const { return chans(type_); }
    size_t elemSize() const { return esz(type_); }
    bool empty() const { return data == nullptr || rows * cols == 0; }
    Size size() const { return Size(cols, rows); }
    bool isContinuous() const { return step == (size_t)cols * elemSize() || rows <= 1; }
    size_t total() const { return (size_t)rows * cols; }

    template <typename T>
    T& at(int i0, int i1) {
        return ((T*)(data + step * i0))[i1];
    }
    template <typename T>
    const T& at(int i0, int i1) const {
        return ((const T*)(data + step * i0))[i1];
    }
    // single index: element of a row or column vector, else raster order (cv::Mat::at(int))
    template <typename T>
    T& at(int i0) {
        if (rows == 1) return ((T*)data)[i0];
        if (cols == 1) return *(T*)(data + step * i0);
        const int i = i0 / cols, j = i0 - i * cols;
        return ((T*)(data + step * i))[j];
    }
    template <typename T>
    const T& at(int i0) const {
        return const_cast<Mat*>(this)->at<T>(i0);
    }

    Mat rowRange(int a, int b) const {
        Mat m(*this);
        m.rows = b - a, m.data = data + step * a;
        return m;
    }
    Mat colRange(int a, int b) const {
        Mat m(*this);
        m.cols = b - a, m.data = data + elemSize() * a;
        return m;
    }
    Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
    Mat clone() const {
        Mat m;
        if (!data) return m;
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memcpy(m.data + m.step * y, data + step * y, (size_t)cols * elemSize());
        return m;
    }
    void copyTo(Mat dst) const {
        if (!dst.data) cvmin_unsupported("Mat::copyTo into an empty header");
        for (int y = 0; y < rows; y++) memcpy(dst.data + dst.step * y, data + step * y, (size_t)cols * elemSize());
    }
    void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const;
    Mat diag() const;
    Mat inv() const;

    static Mat zeros(int r, int c, int type) {
        Mat m(r, c, type);
        memset(m.data, 0, (size_t)r * m.step);
        return m;
    }
    static Mat zeros(Size s, int type) { return zeros(s.height, s.width, type); }
    static Mat ones(int r, int c, int type);
    static Mat ones(Size s, int type) { return ones(s.height, s.width, type); }
    static Mat eye(int r, int c, int type);

    // "m = s": every element becomes s, in place (cv::Mat::operator=(const Scalar&))
    Mat& operator=(const Scalar& s);
};

template <typename T>
struct Mat_ : Mat {
    Mat_(int r, int c) : Mat(r, c, sizeof(T) == 8 ? CV_64F : CV_32F) {}
};

template <typename T>
struct MatCommaInitializer_ {
    Mat m;
    size_t k = 0;
    template <typename U>
    MatCommaInitializer_& operator,(U v) {
        m.at<T>((int)k++) = (T)v;
        return *this;
    }
    operator Mat() const { return m; }
};
template <typename T, typename U>
MatCommaInitializer_<T> operator<<(const Mat_<T>& m, U v) {
    MatCommaInitializer_<T> ci;
    ci.m = m;
    ci.m.template at<T>(0) = (T)v;
    ci.k = 1;
    return ci;
}

// expressions are evaluated eagerly into fresh matrices
Mat operator*(const Mat& a, const Mat& b);  // dense product, CV_32F or CV_64F
Mat operator*(const Mat& a, double s);
Mat operator*(double s, const Mat& a);
Mat operator/(const Mat& a, double s);
Mat operator/(double s, const Mat& a);  // s / a, 0 where a == 0
Mat& operator*=(Mat& a, double s);
Mat& operator/=(Mat& a, double s);
Mat& operator+=(Mat& a, const Mat& b);
inline Mat& operator*=(Mat&& a, double s) { return a *= s; }  // "m.colRange(0, 3) *= s" writes through the view
inline Mat& operator/=(Mat&& a, double s) { return a /= s; }

Scalar sum(const Mat& m);
Scalar mean(const Mat& m);
double norm(const Mat& m, int kind = NORM_L2);
template <typename T, int n>
double norm(const Vec<T, n>& v, int = NORM_L2) {
    double s = 0;
    for (int i = 0; i < n; i++) s += (double)v.val[i] * v.val[i];
    return std::sqrt(s);
}
bool checkRange(const Mat& m);

void Rodrigues(const Mat& R, Vec3f& rvec);
void Rodrigues(const Vec3f& rvec, Mat& R);
void Rodrigues(const Matx33f& R, Vec3f& rvec);

bool eigen(const Matx33f& m, Matx31f& eval, Matx33f& evec);

struct _InputArray {
    _InputArray(const Point3f*, int) {}
    _InputArray(const Point2f*, int) {}
};
bool solvePnP(const _InputArray&, const _InputArray&, const Mat&, const Mat&, Vec3d&, Vec3d&, bool, int);
Mat findEssentialMat(const Mat& pts1, const Mat& pts2, const Mat& K, int method, double prob, double thr, Mat& mask);
int recoverPose(const Mat& E, const Mat& pts1, const Mat& pts2, const Mat& K, Mat& R, Mat& t);

void resize(const Mat& src, Mat& dst, Size, double fx, double fy);
void split(const Mat&, Mat*);
void merge(const std::vector<Mat>&, Mat&);
void cartToPolar(const Mat&, const Mat&, Mat&, Mat&, bool);
void normalize(const Mat&, Mat&, double, double, int);
void cvtColor(const Mat&, Mat&, int);
void imshow(const std::string&, const Mat&);
bool imwrite(const std::string&, const Mat&);
int waitKey(int);
void destroyAllWindows();

}  // namespace cv
