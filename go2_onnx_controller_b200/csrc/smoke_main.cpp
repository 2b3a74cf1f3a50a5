// Smoke executable equivalent to the reference's onnx_inference/src/cpp/main.cpp:24-48:
// zeros[98] in, print model info, time act(), print the 12 actions.  The model path comes from
// argv[1] / $GO2P_MODEL instead of ament_index (ROS 2 is not part of this repository).
#include <array>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <string>

#include "../../include/onnx_actor.hpp"

constexpr unsigned int kInputSize = 98;

static void print_vec(const std::span<float> & vec, const std::string & name)
{
  std::cout << name << ": [";
  for (size_t i = 0; i < vec.size(); ++i) std::cout << vec[i] << (i + 1 < vec.size() ? ",  " : "");
  std::cout << "]" << std::endl;
}

int main(int argc, char ** argv)
{
  const char * env = std::getenv("GO2P_MODEL");
  const std::string model_path = argc > 1 ? argv[1] : (env ? env : "go2_onnx_controller_b200/data/model.onnx");
  std::array<float, kInputSize> observation{};
  std::array<float, 12> action{};
  try {
    ONNXActor actor(model_path, observation, action);
    actor.print_model_info();
    const auto start = std::chrono::steady_clock::now();
    actor.act();   // first call: includes starting the resident kernel (the reference times a cold Run too)
    const auto end = std::chrono::steady_clock::now();
    std::cout << "Inference took " << std::chrono::duration_cast<std::chrono::microseconds>(end - start).count() << "us (cold)" << std::endl;
    const auto s2 = std::chrono::steady_clock::now();
    for (int i = 0; i < 1000; ++i) actor.act();
    const auto e2 = std::chrono::steady_clock::now();
    std::cout << "Warm: " << std::chrono::duration_cast<std::chrono::nanoseconds>(e2 - s2).count() / 1000.0 / 1000.0 << "us per act()" << std::endl;
    print_vec(action, "Action");
  } catch (const std::exception & e) {
    std::cerr << e.what() << std::endl;
    return 1;
  }
  return 0;
}
