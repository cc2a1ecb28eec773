// Stand-in for the ROS 2 generated message header of this name (fields of geometry_msgs/msg/PoseStamped.msg).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's sources).
#pragma once
#include "sensor_msgs/msg/imu.hpp"
namespace geometry_msgs { namespace msg {
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped {
  typedef std::shared_ptr<PoseStamped> SharedPtr;
  std_msgs::msg::Header header;
  Pose pose;
};
} }
