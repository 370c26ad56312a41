// Declaration-only stand-in for ros/ros.h (see tests/stubs/README.md): the logging macros keep printf format checking.
#pragma once
#include <string>
namespace ros {
class NodeHandle {
 public:
  template <class T>
  bool getParam(const std::string& key, T& value) const;
};
}  // namespace ros
extern "C" void uvo_stub_log(const char* fmt, ...) __attribute__((format(printf, 1, 2)));
#define ROS_INFO(...) uvo_stub_log(__VA_ARGS__)
#define ROS_WARN(...) uvo_stub_log(__VA_ARGS__)
#define ROS_ERROR(...) uvo_stub_log(__VA_ARGS__)
#define ROS_FATAL(...) uvo_stub_log(__VA_ARGS__)
