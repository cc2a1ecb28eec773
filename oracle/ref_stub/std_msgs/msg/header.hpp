// Stand-in for the ROS 2 generated message header of this name (plain struct, fields of the .msg definition).
// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's sources).
#pragma once
#include <cstdint>
#include <memory>
#include <string>
namespace builtin_interfaces { namespace msg { struct Time { int32_t sec = 0; uint32_t nanosec = 0; }; } }
namespace std_msgs { namespace msg {
struct Header {
  typedef std::shared_ptr<Header> SharedPtr;
  builtin_interfaces::msg::Time stamp;
  std::string frame_id;
};
} }
