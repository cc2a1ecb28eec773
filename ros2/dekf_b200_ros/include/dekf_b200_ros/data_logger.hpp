// Binary logger writing the on-disk format of the reference's Data_Logger
// (/root/reference/src/decentral_legged_est/include/decentral_legged_est/data_logger.hpp:36-47 file names,
// :49-60,74-85 "<name>,<type>,<length>,\n" rows of <FileName>_Name.csv, :141-187,256-295 one raw record per
// spin_logging() in <FileName>_Data: doubles as 8 bytes, ints as 4-byte floats), so the reference's MATLAB/Python
// post-processing reads the logs of the B200-backed node unchanged.  Own implementation: a list of typed channels
// instead of the reference's pointer/type maps.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

namespace dekf_ros {

class DataLogger {
 public:
  DataLogger() = default;
  ~DataLogger() { close(); }

  // <HOME><location><name>_Data and <HOME><location><name>_Name.csv (reference default location "/log_exp/")
  void init(const std::string &name, const std::string &location = "/log_exp/") {
    const char *home = std::getenv("HOME");
    const std::string base = std::string(home ? home : ".") + location + name;
    data_.open(base + "_Data", std::ios::out | std::ios::binary);
    names_.open(base + "_Name.csv", std::ios::out);
  }
  bool is_open() const { return data_.is_open(); }

  void add_data(const double *p, unsigned length, const std::string &name) {
    channels_.push_back({p, nullptr, length});
    names_ << name << "," << (length == 1 ? "double" : "VectorXd") << "," << length << ",\n";
  }
  void add_data(const std::vector<double> &v, const std::string &name) { add_vector(v.data(), (unsigned)v.size(), name); }
  void add_vector(const double *p, unsigned length, const std::string &name) {
    channels_.push_back({p, nullptr, length});
    names_ << name << ",VectorXd," << length << ",\n";
  }
  void add_data(const int *p, const std::string &name) {
    channels_.push_back({nullptr, p, 1});
    names_ << name << ",int," << 1 << ",\n";
  }

  // one record: every channel in registration order
  void spin_logging() {
    for (const Channel &c : channels_) {
      if (c.d) data_.write(reinterpret_cast<const char *>(c.d), (std::streamsize)(c.length * sizeof(double)));
      else {
        float f = (float)*c.i;
        data_.write(reinterpret_cast<const char *>(&f), sizeof(float));
      }
    }
  }
  void close() {
    if (data_.is_open()) data_.close();
    if (names_.is_open()) names_.close();
  }

 private:
  struct Channel {
    const double *d;
    const int *i;
    unsigned length;
  };
  std::vector<Channel> channels_;
  std::ofstream data_, names_;
};

}  // namespace dekf_ros
