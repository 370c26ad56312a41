// Declaration-only stand-in (see shim/stubs/README.md)
#pragma once
namespace std_msgs {
struct Bool {
  bool data;
};
}  // namespace std_msgs
