// TEST INFRASTRUCTURE -- drives the reference's OWN ONNXController (onnx_controller/src/controller.cpp, compiled from
// where it lies under /root/reference against the stub headers in ref_stubs/) through its public node interface:
// /lowstate and /joy deliveries, robot joint state, the 50 Hz timer callback publish(), and reads back what it
// publishes on /observation_action and hands to Go2RobotInterface::send_command.  Used to pin the restated A1-A6, A9,
// A11 (and the ONNXActor wrapper semantics) to the reference's code; A7 inside the stub ORT is the C restatement.
#include <iostream>
#include <sstream>

#include "controller.hpp"

namespace {
struct Quiet {   // print_vecs() dumps ~260 numbers to stdout per step (controller.cpp:70-154,232)
  std::streambuf* old; std::ostringstream sink;
  Quiet() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Quiet() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

void* refc_create() {
  Quiet q;
  try { return new ONNXController(); } catch (const std::exception& e) { std::cerr << "refc_create: " << e.what() << std::endl; return nullptr; }
}
void refc_destroy(void* h) { delete static_cast<ONNXController*>(h); }

void refc_lowstate(void* h, const float* quat_wxyz, const float* gyro, const float* acc, const int16_t* foot_force) {
  auto m = std::make_shared<unitree_go::msg::LowState>();
  for (int i = 0; i < 4; ++i) { m->imu_state.quaternion[i] = quat_wxyz[i]; m->foot_force[i] = foot_force[i]; }
  for (int i = 0; i < 3; ++i) { m->imu_state.gyroscope[i] = gyro[i]; m->imu_state.accelerometer[i] = acc ? acc[i] : 0.f; }
  static_cast<rclcpp::Node*>(static_cast<ONNXController*>(h))->deliver("/lowstate", m);
}

void refc_joy(void* h, const float* axes, int n_axes, const int32_t* buttons, int n_buttons) {
  auto m = std::make_shared<sensor_msgs::msg::Joy>();
  m->axes.assign(axes, axes + n_axes);
  m->buttons.assign(buttons, buttons + n_buttons);
  static_cast<rclcpp::Node*>(static_cast<ONNXController*>(h))->deliver("/joy", m);
}

void refc_robot(void* h, const double* q, const double* dq, int ready, int safe) {
  Go2RobotInterface* r = static_cast<rclcpp::Node*>(static_cast<ONNXController*>(h))->robot;
  for (int i = 0; i < 12; ++i) { r->q[i] = q[i]; r->dq[i] = dq[i]; }
  r->ready = ready != 0; r->safe = safe != 0;
}

int refc_set_param(void* h, const char* name, double value) {
  return static_cast<rclcpp::Node*>(static_cast<ONNXController*>(h))->set_parameters({rclcpp::Parameter(name, value)}).successful ? 1 : 0;
}

// one 50 Hz tick.  Returns 1 if publish() ran to the end (a command was sent), 0 if it was gated off.
int refc_step(void* h, float* obs98, float* action12, double* q_des12, double* kp12, double* kd12) {
  rclcpp::Node* n = static_cast<rclcpp::Node*>(static_cast<ONNXController*>(h));
  const int before = n->robot->n_commands;
  { Quiet q; n->fire_timer(); }
  if (n->robot->n_commands == before) return 0;
  const auto msg = std::any_cast<onnx_interfaces::msg::ObservationAction>(n->last_published.at("/observation_action"));
  for (int i = 0; i < 98; ++i) obs98[i] = msg.observation[i];
  for (int i = 0; i < 12; ++i) {
    action12[i] = msg.action[i];
    q_des12[i] = n->robot->cmd_q[i]; kp12[i] = n->robot->cmd_kp[i]; kd12[i] = n->robot->cmd_kd[i];
  }
  return 1;
}

}  // extern "C"
