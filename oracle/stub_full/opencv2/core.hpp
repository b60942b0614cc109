// Declaration-only stand-in for the OpenCV API surface the reference's entry points use (cnn.cpp, inference.cpp,
// grad_cam.cpp, pipeline.cpp: imread / resize / imshow / flip / warpAffine / applyColorMap / Mat arithmetic).
// TEST INFRASTRUCTURE: it exists so that `make -C cnn_b200/host refcheck` can run those UNMODIFIED sources through
// `g++ -fsyntax-only` against the B200 layer headers on a machine without OpenCV -- nothing links against it.
#pragma once
#include <algorithm>
#include <cassert>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <memory>
#include <random>
#include <sstream>
#include <string>
#include <tuple>
#include <vector>

typedef unsigned char uchar;
#define CV_8UC1 0
#define CV_8UC3 16
#define CV_32FC3 21
#define CV_VERSION "syntax-check stand-in"

namespace cv {
template <class T> struct Size_ {
    T width{}, height{};
    Size_() = default;
    Size_(T w, T h) : width(w), height(h) {}
    template <class U> Size_(const Size_<U>& o) : width((T)o.width), height((T)o.height) {}
};
using Size = Size_<int>;
using Size2f = Size_<float>;
template <class T> struct Point_ {
    T x{}, y{};
    Point_() = default;
    Point_(T a, T b) : x(a), y(b) {}
};
using Point2f = Point_<float>;
template <class T> struct Rect_ {
    T x{}, y{}, width{}, height{};
    Rect_() = default;
    Rect_(T a, T b, T w, T h) : x(a), y(b), width(w), height(h) {}
    Size_<T> size() const;
};
using Rect = Rect_<int>;
using Rect2f = Rect_<float>;

struct Mat {
    int rows = 0, cols = 0;
    uchar* data = nullptr;
    Mat();
    Mat(int r, int c, int type);
    bool empty() const;
    Size size() const;
    Mat clone() const;
    Mat operator()(const Rect& roi) const;
    void convertTo(Mat& dst, int type) const;
    template <class T> T& at(int r, int c);
    template <class T> T* begin();
    template <class T> T* end();
};
Mat operator-(int a, const Mat& b);
Mat operator+(const Mat& a, const Mat& b);
Mat operator/(const Mat& a, double b);
Mat operator*(const Mat& a, double b);

struct RotatedRect {
    RotatedRect(const Point2f& center, const Size2f& size, float angle);
    Rect2f boundingRect2f() const;
};
template <class T> T saturate_cast(float v);
void flip(const Mat& src, Mat& dst, int code);
}  // namespace cv
