// Stand-in for the generated header of the reference's src/communication/optitrack_broadcast/msg/Mocap.msg.
// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's sources).
#pragma once
#include <array>
#include "std_msgs/msg/header.hpp"
namespace optitrack_broadcast { namespace msg {
struct Mocap {
  typedef std::shared_ptr<Mocap> SharedPtr;
  std_msgs::msg::Header header;
  std::array<float, 3> position{}, velocity{}, angular_velocity{};
  std::array<float, 4> quaternion{};
};
} }
