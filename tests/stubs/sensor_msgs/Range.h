// Declaration-only stand-in (see tests/stubs/README.md)
#pragma once
namespace sensor_msgs {
struct Range {};
}  // namespace sensor_msgs
