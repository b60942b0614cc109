// Minimal stand-in for <opencv2/core.hpp>, used ONLY to compile the reference's
// layer sources (which use OpenCV for image I/O types, never for arithmetic)
// into oracle/_ref without an OpenCV install.  Test infrastructure.
// It also pulls in the std headers the reference relies on OpenCV to include.
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
#define CV_VERSION "stub"

namespace cv {
struct Mat {
    int rows = 0, cols = 0, kind = 0;
    std::shared_ptr<std::vector<uchar>> store;
    uchar* data = nullptr;
    Mat() = default;
    Mat(int r, int c, int t)
        : rows(r), cols(c), kind(t),
          store(std::make_shared<std::vector<uchar>>(size_t(r) * c * (t == CV_8UC3 ? 3 : 1))),
          data(store->data()) {}
    bool empty() const { return data == nullptr; }
};
template <class T>
inline T saturate_cast(float v) {  // cvRound (ties-to-even) then clamp
    const long r = std::lrintf(v);
    return T(r < 0 ? 0 : (r > 255 ? 255 : r));
}
}  // namespace cv
