// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "rclcpp/rclcpp.hpp"
namespace message_filters {
// the node hands over std::shared_ptr<rclcpp::Node>(this) (stereo-pub-node.cpp:72-73): like the real class the stand-in keeps
// that handle -- for the life of the process, so that the temporary owner never deletes the node
inline std::vector<std::shared_ptr<rclcpp::Node>> &kept_nodes() { static auto *v = new std::vector<std::shared_ptr<rclcpp::Node>>(); return *v; }
template <typename M>
class Subscriber {
 public:
  Subscriber(std::shared_ptr<rclcpp::Node> node, const std::string &topic) : topic_(topic) { kept_nodes().push_back(node); }
  const std::string &topic() const { return topic_; }
 private:
  std::string topic_;
};
}  // namespace message_filters
