// declaration-only stand-in, see core.hpp
#pragma once
#include "core.hpp"
namespace cv {
enum { COLORMAP_JET = 2 };
void resize(const Mat& src, Mat& dst, Size size);
void warpAffine(const Mat& src, Mat& dst, const Mat& m, Size size);
Mat getRotationMatrix2D(Point2f center, double angle, double scale);
void applyColorMap(const Mat& src, Mat& dst, int colormap);
}  // namespace cv
