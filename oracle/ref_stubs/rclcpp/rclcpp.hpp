// TEST INFRASTRUCTURE -- minimal stand-in for the rclcpp surface ONNXController touches (controller.cpp:20-62,
// 254-285): subscriptions, a publisher, parameters, a wall timer, a logger.  The harness drives it: deliver() hands a
// message to the subscription callback of a topic, fire_timer() runs the 50 Hz callback (publish()).
#pragma once
#include <algorithm>
#include <any>
#include <array>
#include <cstdlib>
#include <iostream>
#include <string_view>
#include <chrono>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace rcl_interfaces { namespace msg { struct SetParametersResult { bool successful = true; std::string reason; }; } }

class Go2RobotInterface;

namespace rclcpp {
struct Logger {};
struct TimerBase { using SharedPtr = std::shared_ptr<TimerBase>; };
template <class M> struct Subscription { using SharedPtr = std::shared_ptr<Subscription<M>>; };
struct Parameter {
  std::string name; double value;
  Parameter(std::string n, double v) : name(std::move(n)), value(v) {}
  const std::string& get_name() const { return name; }
  double as_double() const { return value; }
};
namespace node_interfaces { struct OnSetParametersCallbackHandle { using SharedPtr = std::shared_ptr<OnSetParametersCallbackHandle>; }; }

class Node;
template <class M> struct Publisher {
  using SharedPtr = std::shared_ptr<Publisher<M>>;
  Node* node; std::string topic;
  void publish(const M& m);
};

class Node {
public:
  explicit Node(const std::string& name) : name_(name) {}
  virtual ~Node() = default;
  template <class M, class CB>
  typename Subscription<M>::SharedPtr create_subscription(const std::string& topic, int, CB cb) {
    std::function<void(typename M::SharedPtr)> f = cb;
    subs_[topic] = [f](std::any a) { f(std::any_cast<typename M::SharedPtr>(a)); };
    return std::make_shared<Subscription<M>>();
  }
  template <class M>
  typename Publisher<M>::SharedPtr create_publisher(const std::string& topic, int) {
    auto p = std::make_shared<Publisher<M>>(); p->node = this; p->topic = topic; return p;
  }
  template <class T> void declare_parameter(const std::string& n, T v) { params_[n] = (double)v; }
  template <class CB>
  node_interfaces::OnSetParametersCallbackHandle::SharedPtr add_on_set_parameters_callback(CB cb) {
    param_cb_ = cb; return std::make_shared<node_interfaces::OnSetParametersCallbackHandle>();
  }
  template <class D, class CB> TimerBase::SharedPtr create_wall_timer(D, CB cb) { timer_cb_ = cb; return std::make_shared<TimerBase>(); }
  Logger get_logger() const { return {}; }

  // ---- harness side
  template <class M> void deliver(const std::string& topic, std::shared_ptr<M> msg) { subs_.at(topic)(std::any(msg)); }
  void fire_timer() { timer_cb_(); }
  rcl_interfaces::msg::SetParametersResult set_parameters(const std::vector<Parameter>& p) { return param_cb_(p); }
  std::map<std::string, std::any> last_published;
  std::map<std::string, int> publish_count;
  Go2RobotInterface* robot = nullptr;

private:
  std::string name_;
  std::map<std::string, std::function<void(std::any)>> subs_;
  std::map<std::string, double> params_;
  std::function<rcl_interfaces::msg::SetParametersResult(const std::vector<Parameter>&)> param_cb_;
  std::function<void()> timer_cb_;
};

template <class M> void Publisher<M>::publish(const M& m) { node->last_published[topic] = m; node->publish_count[topic]++; }

inline void init(int, char**) {}
inline void spin(std::shared_ptr<Node>) {}
inline void shutdown() {}
}  // namespace rclcpp

#define RCLCPP_INFO(logger, ...) do { (void)(logger); } while (0)
#define RCLCPP_WARN(logger, ...) do { (void)(logger); } while (0)
