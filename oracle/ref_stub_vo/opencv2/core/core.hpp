// TEST INFRASTRUCTURE ONLY: stand-in header written for this repository so that the reference's VO wrapper node
// (visual_odometry/orbslam3_ros2/src/stereo-decentralized/stereo-pub-node.cpp) compiles UNMODIFIED into oracle/_ref/vo_pin
// (recipe: oracle/Makefile vo_pin; driver: oracle/vo_pin_main.cc).  ORB-SLAM3, OpenCV, Sophus, cv_bridge and message_filters are
// absent from the image and out of scope; only the surface that source file touches exists here.
#pragma once
#include <string>
#define CV_32F 5
namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c) : rows(r), cols(c) {}
  bool empty() const { return rows == 0 || cols == 0; }
  Mat clone() const { return *this; }
  Mat rowRange(int a, int b) const { return Mat(b - a, cols); }
  Mat colRange(int a, int b) const { return Mat(rows, b - a); }
};
class FileNode {
 public:
  operator int() const { return 0; }
  void operator>>(Mat &) const {}
};
class FileStorage {
 public:
  enum { READ = 0 };
  FileStorage(const std::string &, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const char *) const { return FileNode(); }
};
inline void resize(const Mat &src, Mat &dst, Size s) { dst = Mat(s.height, s.width); (void)src; }
inline void initUndistortRectifyMap(const Mat &, const Mat &, const Mat &, const Mat &, Size, int, Mat &, Mat &) {}
}  // namespace cv
