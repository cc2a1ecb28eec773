// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
//
// ORB_SLAM3::System: TrackStereo() returns the next pose of a script the driver sets (refstub_vo::script()); the real header
// is also where the node gets its unqualified std names from (using namespace std).
#pragma once
#include <cmath>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "opencv2/core/core.hpp"
#include "sophus/se3.hpp"
using namespace std;
namespace refstub_vo {
struct Script { std::vector<Sophus::SE3f> poses; size_t next = 0; std::vector<double> stamps_seen; float image_scale = 1.f; bool shut = false; };
inline Script &script() { static Script s; return s; }
}  // namespace refstub_vo
namespace ORB_SLAM3 {
class System {
 public:
  float GetImageScale() { return refstub_vo::script().image_scale; }
  Sophus::SE3f TrackStereo(const cv::Mat &, const cv::Mat &, const double &timestamp) {
    refstub_vo::Script &s = refstub_vo::script();
    s.stamps_seen.push_back(timestamp);
    return s.poses.at(s.next++);
  }
  void Shutdown() { refstub_vo::script().shut = true; }
};
}  // namespace ORB_SLAM3
