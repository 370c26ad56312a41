// Declaration-only stand-in (see tests/stubs/README.md): VO_utility.h says `using namespace cv::xfeatures2d;`
#pragma once
#include <opencv2/opencv.hpp>
namespace cv {
namespace xfeatures2d {}
}  // namespace cv
