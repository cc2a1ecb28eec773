// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
#pragma once
#include <memory>
#include <stdexcept>
#include "opencv2/core/core.hpp"
#include "sensor_msgs/msg/image.hpp"
namespace cv_bridge {
class Exception : public std::runtime_error {
 public:
  explicit Exception(const std::string &w) : std::runtime_error(w) {}
};
struct CvImage { std_msgs::msg::Header header; std::string encoding; cv::Mat image; };
typedef std::shared_ptr<CvImage> CvImagePtr;
typedef std::shared_ptr<const CvImage> CvImageConstPtr;
inline CvImageConstPtr toCvShare(const sensor_msgs::msg::Image::SharedPtr &m) {
  auto p = std::make_shared<CvImage>();
  p->header = m->header;
  p->image = cv::Mat((int)m->height, (int)m->width);
  return p;
}
}  // namespace cv_bridge
