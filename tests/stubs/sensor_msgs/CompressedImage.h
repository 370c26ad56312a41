// Declaration-only stand-in (see tests/stubs/README.md)
#pragma once
#include <memory>
#include <string>
#include <vector>
namespace sensor_msgs {
struct CompressedImage {
  std::string format;
  std::vector<unsigned char> data;
  typedef std::shared_ptr<const CompressedImage> ConstPtr;
};
}  // namespace sensor_msgs
