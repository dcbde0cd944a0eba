// TEST INFRASTRUCTURE (oracle). Not part of the shipped product.
//
// Declaration-only stand-in for the three OpenCV headers that the reference's voldor/utils.h includes, just enough for
// the reference's voldor/config.h (which includes utils.h) to be PARSED unmodified by g++: oracle/ref_shim/
// ref_config_probe.cpp uses the reference's own `struct Config` — defaults and flag parser — as the oracle for
// csrc/config.h.  Nothing declared here is ever defined or called (OpenCV's C++ library is not in this image).
#pragma once
#include <cstdio>
#include <ostream>
#include <string>
#include <typeinfo>
namespace cv {
enum { CV_32F = 5, NORM_L2 = 4 };
struct Scalar {
    double operator[](int) const;
};
template <typename T, int n>
struct Vec {
    T val[n];
    Vec();
    Vec(T, T, T);
    Vec(T, T, T, T, T, T);
    T& operator[](int);
};
typedef Vec<float, 2> Vec2f;
typedef Vec<float, 3> Vec3f;
typedef Vec<float, 4> Vec4f;
typedef Vec<float, 6> Vec6f;
typedef Vec<double, 3> Vec3d;
template <typename T, int n>
std::ostream& operator<<(std::ostream&, const Vec<T, n>&);
struct Mat {
    Mat();
    static Mat eye(int, int, int);
    static Mat zeros(int, int, int);
    template <typename T>
    T& at(int);
    Mat diag() const;
};
void Rodrigues(const Mat&, Vec3f&);
Scalar mean(const Mat&);
double norm(const Mat&, int);
double norm(const Vec3f&, int);
}  // namespace cv
