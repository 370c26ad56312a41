// Declaration-only stand-in (see tests/stubs/README.md)
#pragma once
namespace std_msgs {
struct Bool {
  bool data;
};
}  // namespace std_msgs
