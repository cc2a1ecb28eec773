// Stand-in for the ROS 2 generated message header of this name (plain struct, fields of sensor_msgs/msg/Imu.msg).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's sources).
#pragma once
#include <array>
#include "std_msgs/msg/header.hpp"
namespace geometry_msgs { namespace msg {
#ifndef REFSTUB_GEOMETRY_BASICS
#define REFSTUB_GEOMETRY_BASICS
struct Vector3 { double x = 0, y = 0, z = 0; };
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
#endif
} }
namespace sensor_msgs { namespace msg {
struct Imu {
  typedef std::shared_ptr<Imu> SharedPtr;
  std_msgs::msg::Header header;
  geometry_msgs::msg::Quaternion orientation;
  std::array<double, 9> orientation_covariance{};
  geometry_msgs::msg::Vector3 angular_velocity;
  std::array<double, 9> angular_velocity_covariance{};
  geometry_msgs::msg::Vector3 linear_acceleration;
  std::array<double, 9> linear_acceleration_covariance{};
};
} }
