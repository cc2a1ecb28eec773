// Stand-in for the ROS 2 generated message header of this name (fields of sensor_msgs/msg/JointState.msg).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's sources).
#pragma once
#include <string>
#include <vector>
#include "std_msgs/msg/header.hpp"
namespace sensor_msgs { namespace msg {
struct JointState {
  typedef std::shared_ptr<JointState> SharedPtr;
  std_msgs::msg::Header header;
  std::vector<std::string> name;
  std::vector<double> position, velocity, effort;
};
} }
