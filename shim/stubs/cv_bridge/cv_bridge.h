// Declaration-only stand-in (see shim/stubs/README.md)
#pragma once
#include "sensor_msgs/CompressedImage.h"
