#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace sensor_msgs { namespace msg { struct Joy { using SharedPtr = std::shared_ptr<Joy>; std::vector<float> axes; std::vector<int32_t> buttons; }; } }
