// Declaration-only stand-in (see tests/stubs/README.md): math_utility.h names geometry_msgs::Vector3
#pragma once
namespace geometry_msgs {
struct Vector3 {
  double x, y, z;
};
}  // namespace geometry_msgs
