// TEST INFRASTRUCTURE -- stand-in for the out-of-tree go2_control_interface (README.md:5): producer of q/dq and
// consumer of the command, as used at controller.cpp:27,58,158-171,187-191,251.
#pragma once
#include <array>
#include <string_view>

#include "mini_eigen.hpp"
#include "rclcpp/rclcpp.hpp"

class Go2RobotInterface {
public:
  using Vec12 = Eigen::Vector<double, 12>;
  Go2RobotInterface(rclcpp::Node& node, const std::array<std::string_view, 12>& names) : names_(names) { node.robot = this; }
  bool is_ready() const { return ready; }
  bool is_safe() const { return safe; }
  const Vec12& get_q() const { return q; }
  const Vec12& get_dq() const { return dq; }
  void start_async(const Vec12& q0) { q_start = q0; started = true; }
  void send_command(const Vec12& q_des, const Vec12& dq_des, const Vec12& tau, const Vec12& kp, const Vec12& kd) {
    cmd_q = q_des; cmd_dq = dq_des; cmd_tau = tau; cmd_kp = kp; cmd_kd = kd; ++n_commands;
  }
  // harness side
  Vec12 q, dq, q_start, cmd_q, cmd_dq, cmd_tau, cmd_kp, cmd_kd;
  bool ready = true, safe = true, started = false;
  int n_commands = 0;
private:
  std::array<std::string_view, 12> names_;
};
