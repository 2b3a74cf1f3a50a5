#pragma once
#include <array>
#include <cstdint>
#include <memory>
namespace unitree_go { namespace msg {
struct IMUState { std::array<float, 4> quaternion{}; std::array<float, 3> gyroscope{}; std::array<float, 3> accelerometer{}; };
struct LowState { using SharedPtr = std::shared_ptr<LowState>; IMUState imu_state; std::array<int16_t, 4> foot_force{}; };
} }
