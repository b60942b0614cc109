// declaration-only stand-in, see core.hpp
#pragma once
#include "core.hpp"
namespace cv {
enum { IMWRITE_PNG_COMPRESSION = 16 };
Mat imread(const std::string& path);
bool imwrite(const std::string& path, const Mat& img, const std::vector<int>& params = std::vector<int>());
void imshow(const std::string& name, const Mat& img);
int waitKey(int delay = 0);
void destroyAllWindows();
}  // namespace cv
