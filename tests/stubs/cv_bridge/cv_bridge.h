// Declaration-only stand-in (see tests/stubs/README.md)
#pragma once
#include "sensor_msgs/CompressedImage.h"
