// Stand-in for <rclcpp/rclcpp.hpp> (ROS 2 is absent from the image).
//
// TEST INFRASTRUCTURE ONLY: a deterministic, single-threaded replacement of the few rclcpp facilities the reference's
// node classes use (orien_ekf.cpp:8-45, EstSub.cpp:7-28,110-209, go1Sub.cpp:8-24), so that those classes compile
// UNMODIFIED into oracle/_ref/ and can be driven in lock-step by oracle/ref_nodes.cc:
//   * the clock is a global the driver sets (refstub::now_ns());
//   * publishers deliver synchronously to every subscription of the same topic in this process;
//   * wall timers never fire by themselves: the driver calls Node::fire_timers();
//   * declare_parameter() takes the value from refstub::param_overrides()["<node>.<param>"] when present, else the
//     default given by the node -- the role of the YAML file in the reference's launch file.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "std_msgs/msg/header.hpp"

#ifndef RCLCPP_ERROR
#define RCLCPP_ERROR(logger, ...) do { (void)(logger); std::fprintf(stderr, __VA_ARGS__); std::fputc('\n', stderr); } while (0)
#endif

namespace refstub {
struct ParamValue {
  int kind = 0;  // 1 int, 2 double, 3 bool, 4 string, 5 double array
  int64_t i = 0;
  double d = 0.0;
  bool b = false;
  std::string s;
  std::vector<double> v;
};
inline int64_t &now_ns() { static int64_t t = 0; return t; }
inline std::map<std::string, ParamValue> &param_overrides() { static std::map<std::string, ParamValue> m; return m; }
inline std::string topic_key(const std::string &t) { return (!t.empty() && t[0] == '/') ? t.substr(1) : t; }
template <typename Msg>
std::map<std::string, std::vector<std::function<void(std::shared_ptr<Msg>)>>> &bus() {
  static std::map<std::string, std::vector<std::function<void(std::shared_ptr<Msg>)>>> b;
  return b;
}
inline std::vector<std::function<void()>> &bus_clearers() { static std::vector<std::function<void()>> v; return v; }
inline void clear_bus() { for (auto &f : bus_clearers()) f(); }
template <typename Msg>
void deliver(const std::string &topic, const Msg &m) {
  auto it = bus<Msg>().find(topic_key(topic));
  if (it == bus<Msg>().end()) return;
  for (auto &cb : it->second) cb(std::make_shared<Msg>(m));
}
}  // namespace refstub

namespace rclcpp {

class Time {
 public:
  Time() : ns_(0) {}
  explicit Time(int64_t ns) : ns_(ns) {}
  int64_t nanoseconds() const { return ns_; }
  double seconds() const { return (double)ns_ / 1e9; }
  operator builtin_interfaces::msg::Time() const {  // header.stamp = node->now()
    builtin_interfaces::msg::Time t;
    t.sec = (int32_t)(ns_ / 1000000000LL);
    t.nanosec = (uint32_t)(ns_ % 1000000000LL);
    return t;
  }
 private:
  int64_t ns_;
};
class Clock {
 public:
  Time now() const { return Time(refstub::now_ns()); }
};

class Parameter {
 public:
  Parameter() {}
  explicit Parameter(const refstub::ParamValue &v) : v_(v) {}
  int64_t as_int() const { return v_.i; }
  double as_double() const { return v_.d; }
  bool as_bool() const { return v_.b; }
  std::string as_string() const { return v_.s; }
  std::vector<double> as_double_array() const { return v_.v; }
 private:
  refstub::ParamValue v_;
};

template <typename Msg>
class Subscription {
 public:
  typedef std::shared_ptr<Subscription<Msg>> SharedPtr;
};
template <typename Msg>
class Publisher {
 public:
  typedef std::shared_ptr<Publisher<Msg>> SharedPtr;
  explicit Publisher(const std::string &topic) : topic_(topic) {}
  void publish(const Msg &m) { refstub::deliver<Msg>(topic_, m); }
 private:
  std::string topic_;
};
class TimerBase {
 public:
  typedef std::shared_ptr<TimerBase> SharedPtr;
  std::function<void()> callback;
};
class Logger {};

class Node {
 public:
  explicit Node(const std::string &name) : name_(name) {}
  virtual ~Node() {}
  const char *get_name() const { return name_.c_str(); }
  Logger get_logger() const { return Logger(); }

  template <typename T> void declare_parameter(const std::string &n, const T &def) { store(n, make_value(def)); }
  void declare_parameter(const std::string &n, const char *def) { store(n, make_value(std::string(def))); }
  Parameter get_parameter(const std::string &n) const { return Parameter(params_.at(n)); }

  template <typename Msg, typename Cb>
  typename Subscription<Msg>::SharedPtr create_subscription(const std::string &topic, int /*qos*/, Cb cb) {
    static bool registered = false;
    if (!registered) { registered = true; refstub::bus_clearers().push_back([] { refstub::bus<Msg>().clear(); }); }
    refstub::bus<Msg>()[refstub::topic_key(topic)].push_back(std::function<void(std::shared_ptr<Msg>)>(cb));
    return std::make_shared<Subscription<Msg>>();
  }
  template <typename Msg>
  typename Publisher<Msg>::SharedPtr create_publisher(const std::string &topic, int /*qos*/) {
    return std::make_shared<Publisher<Msg>>(topic);
  }
  template <typename Dur, typename Cb>
  TimerBase::SharedPtr create_wall_timer(Dur, Cb cb) {
    auto t = std::make_shared<TimerBase>();
    t->callback = cb;
    timers_.push_back(t);
    return t;
  }
  void fire_timers() { for (auto &t : timers_) t->callback(); }

 private:
  static refstub::ParamValue make_value(const std::vector<double> &d) { refstub::ParamValue p; p.kind = 5; p.v = d; return p; }
  static refstub::ParamValue make_value(int d) { refstub::ParamValue p; p.kind = 1; p.i = d; return p; }
  static refstub::ParamValue make_value(double d) { refstub::ParamValue p; p.kind = 2; p.d = d; return p; }
  static refstub::ParamValue make_value(bool d) { refstub::ParamValue p; p.kind = 3; p.b = d; return p; }
  static refstub::ParamValue make_value(const std::string &d) { refstub::ParamValue p; p.kind = 4; p.s = d; return p; }
  void store(const std::string &n, refstub::ParamValue def) {
    auto it = refstub::param_overrides().find(name_ + "." + n);
    params_[n] = (it != refstub::param_overrides().end()) ? it->second : def;
  }
  std::string name_;
  std::map<std::string, refstub::ParamValue> params_;
  std::vector<TimerBase::SharedPtr> timers_;
};

inline void init(int, char **) {}
inline void init(int, char *const *) {}
inline void shutdown() {}
inline void spin(std::shared_ptr<Node>) {}
inline bool ok() { return true; }

}  // namespace rclcpp
