// Exact operation tally of the committed per-instance algorithm (csrc/estimator_core.cuh), obtained by
// instantiating the kernel bodies with a counting scalar.  Dev tool: prints the numbers quoted in DESIGN.md
// and used by bench.py's roofline (tests/test_flop_tally.py keeps them in sync).  Not a product path.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

struct Cnt {
  static long add, mul, div, sq, trig;
  static void reset() { add = mul = div = sq = trig = 0; }
  static long flops() { return add + mul + div + sq; }
};
long Cnt::add = 0, Cnt::mul = 0, Cnt::div = 0, Cnt::sq = 0, Cnt::trig = 0;

struct CT {
  double v;
  CT() : v(0) {}
  CT(double x) : v(x) {}
  CT(int x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
};
inline CT operator+(CT a, CT b) { Cnt::add++; return CT(a.v + b.v); }
inline CT operator-(CT a, CT b) { Cnt::add++; return CT(a.v - b.v); }
inline CT operator*(CT a, CT b) { Cnt::mul++; return CT(a.v * b.v); }
inline CT operator/(CT a, CT b) { Cnt::div++; return CT(a.v / b.v); }
inline CT operator-(CT a) { return CT(-a.v); }
inline CT &operator+=(CT &a, CT b) { Cnt::add++; a.v += b.v; return a; }
inline CT &operator-=(CT &a, CT b) { Cnt::add++; a.v -= b.v; return a; }
inline CT &operator*=(CT &a, CT b) { Cnt::mul++; a.v *= b.v; return a; }
inline bool operator!=(CT a, CT b) { return a.v != b.v; }
inline bool operator==(CT a, CT b) { return a.v == b.v; }
inline CT sqrt(CT a) { Cnt::sq++; return CT(std::sqrt(a.v)); }

#include "../../decentralized_ekf_mhe_b200/csrc/host_setup.hpp"
namespace dekf {
template <>
inline void sincos_t<CT>(CT x, CT *s, CT *c) {
  Cnt::trig++;
  *s = CT(std::sin(x.v));
  *c = CT(std::cos(x.v));
}
}  // namespace dekf
using namespace dekf;

int main() {
  dekf_config cfg;
  fill_go1_defaults(&cfg);
  cfg.ekf_rate = 200;
  cfg.n_instances = 1;
  MheConst<CT> mc = make_mhe_const<CT>(cfg);
  EkfConst<CT> ec = make_ekf_const<CT>(cfg);
  // a generic stage
  Cov9<CT> P;
  Vec9<CT> x;
  for (int f = 0; f < 6; ++f) P.pp.a[f] = P.vv.a[f] = P.bb.a[f] = CT(f == 0 || f == 3 || f == 5 ? 1e-3 : 1e-5);
  for (int f = 0; f < 9; ++f) P.pv.a[f] = P.pb.a[f] = P.vb.a[f] = CT(1e-6 * (f + 1));
  x.p = x.v = x.b = v3<CT>(CT(0.1), CT(0.2), CT(0.3));
  S3<CT> Lam;
  for (int f = 0; f < 6; ++f) Lam.a[f] = CT(f == 0 || f == 3 || f == 5 ? 200.0 : 3.0);
  V3<CT> eta = v3<CT>(CT(1.0), CT(2.0), CT(3.0));
  M3<CT> R = quat_to_rot<CT>(CT(0.9), CT(0.1), CT(0.2), CT(0.3));
  V3<CT> as = v3<CT>(CT(0.1), CT(0.0), CT(0.2)), dlt = v3<CT>(CT(0.002), CT(0.0), CT(0.0));
  // the committed sweep: DefaultMath = version-4 measurement stage + version-5 propagation; PPM 0 = the loop body of the
  // full re-sweep (P_pp not carried), PPM 1 = the peeled arrival stage / incremental sweep / KF alternative
  using Mth = DefaultMath<CT>;
  Cnt::reset();
  Mth::meas<0>(P, x, Lam, eta);
  long f_meas = Cnt::flops(), d_meas = Cnt::div;
  Cnt::reset();
  Mth::prop<0>(mc, P, x, R, as, false, dlt);
  long f_prop = Cnt::flops();
  Cnt::reset();
  Mth::prop<0>(mc, P, x, R, as, true, dlt);
  long f_prop_vo = Cnt::flops();
  Cnt::reset();
  Mth::meas<1>(P, x, Lam, eta);
  long f_meas_pp = Cnt::flops();
  Cnt::reset();
  Mth::prop<1>(mc, P, x, R, as, false, dlt);
  long f_prop_pp = Cnt::flops();
  Cnt::reset();
  Mth::prop<1>(mc, P, x, R, as, true, dlt);
  long f_prop_vo_pp = Cnt::flops();
  // EKF
  EkfState<CT> s;
  for (int f = 0; f < 4; ++f) s.q[f] = CT(f == 0 ? 1.0 : 0.01);
  for (int f = 0; f < 16; ++f) s.P[f] = CT(f % 5 == 0 ? 1e-6 : 1e-9);
  CT w[3] = {CT(0.1), CT(0.2), CT(0.3)}, a[3] = {CT(0.1), CT(0.2), CT(9.8)}, qv[4] = {CT(1.0), CT(0.0), CT(0.0), CT(0.0)};
  Cnt::reset();
  ekf_predict(ec, s, w);
  long f_pred = Cnt::flops();
  Cnt::reset();
  ekf_correct(ec, s, a);
  long f_corr = Cnt::flops();
  Cnt::reset();
  ekf_vo_correct(ec, s, qv);
  long f_vo = Cnt::flops();
  // assembly (one instance, all four legs in stance = the most expensive branch-free path)
  Dims dm = make_dims(cfg);
  StateSizes ss = state_sizes(dm);
  std::vector<CT> win(ss.win), dummyT(64);
  std::vector<double> d(4096, 0.0);
  std::vector<uint8_t> fl(64, 0);
  std::vector<int32_t> i32(64, 0);
  Buffers<CT> b;
  std::memset(&b, 0, sizeof(b));
  b.win = win.data();
  b.hist_time = d.data();
  b.hist_quat = d.data() + 128;
  b.wp = d.data() + 1024;
  b.wp_time = d.data() + 1100;
  b.wp_count = i32.data();
  b.p_vo = d.data() + 1200;
  b.pend_flag = fl.data() + 32;
  b.pend = d.data() + 1300;
  b.status = i32.data() + 8;
  std::vector<double> in_d(64, 0.3);
  for (int k = 0; k < 4; ++k) in_d[40 + k] = 200.0;
  Inputs in;
  std::memset(&in, 0, sizeof(in));
  in.gyro = in_d.data();
  in.accel = in_d.data() + 3;
  in.imu_time = in_d.data() + 6;
  in.joint_pos = in_d.data() + 8;
  in.joint_vel = in_d.data() + 20;
  in.foot_force = in_d.data() + 40;
  Outputs out;
  std::memset(&out, 0, sizeof(out));
  const double q[4] = {0.9, 0.1, 0.2, 0.3};
  Cnt::reset();
  mhe_assemble<CT, Go1Model<CT>>(mc, dm, b, in, out, 1, 0, q);
  long f_asm = Cnt::flops(), t_asm = Cnt::trig;
  std::printf("{\"meas_update\": %ld, \"propagate\": %ld, \"propagate_vo\": %ld, \"ekf_predict\": %ld, "
              "\"ekf_correct\": %ld, \"ekf_vo_correct\": %ld, \"assemble_go1\": %ld, \"assemble_go1_sincos\": %ld, "
              "\"meas_update_div\": %ld, \"meas_update_pp\": %ld, \"propagate_pp\": %ld, \"propagate_vo_pp\": %ld}\n",
              f_meas, f_prop, f_prop_vo, f_pred, f_corr, f_vo, f_asm, t_asm, d_meas, f_meas_pp, f_prop_pp, f_prop_vo_pp);
  return 0;
}
