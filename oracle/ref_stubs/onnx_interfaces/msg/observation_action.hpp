// layout of onnx_interfaces/msg/ObservationAction.msg:1-2 (float32[98] observation, float32[12] action)
#pragma once
#include <array>
#include <memory>
namespace onnx_interfaces { namespace msg {
struct ObservationAction { using SharedPtr = std::shared_ptr<ObservationAction>; std::array<float, 98> observation{}; std::array<float, 12> action{}; };
} }
