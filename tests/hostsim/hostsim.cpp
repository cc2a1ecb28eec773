// MATH-DEBUG HARNESS, NOT A PRODUCT PATH.
// Compiles the per-instance kernel bodies (csrc/estimator_core.cuh, __host__ __device__) with g++
// and loops them over instances on the CPU so that the algebra of the kernels can be checked against
// the oracle in the GPU-less build container.  It is built only by tests/test_hostsim.py into
// tests/hostsim/_build/, is not part of libdekf_b200.so, is not importable from the package and is
// never timed.  The product path has no CPU fallback (dekf_create fails without a CUDA device).
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../decentralized_ekf_mhe_b200/csrc/host_setup.hpp"

using namespace dekf;

namespace {
// general linear rows for the NEXT hostsim_run (hostsim_set_rows): what dekf_add_state_rows does on a handle
int g_rows = 0;
double g_row_a[81], g_row_lb[9], g_row_ub[9];

template <typename T>
struct Sim {
  double Vrow[81];
  dekf_config cfg;
  Dims dm;
  EkfConst<T> ec;
  MheConst<T> mc;
  BoxConst bc;
  BoxBuffers bb;
  FootConst fc;
  FootBuffers fb;
  Buffers<T> b;
  std::vector<std::vector<char>> store;

  template <typename U>
  U *alloc(size_t count) {
    store.emplace_back(count * sizeof(U) + 64, 0);
    return reinterpret_cast<U *>(store.back().data());
  }
  explicit Sim(const dekf_config &c) : cfg(c) {
    dm = make_dims(c);
    ec = make_ekf_const<T>(c);
    mc = make_mhe_const<T>(c);
    bc = make_box_const(c);
    bb.fac = alloc<double>((size_t)dm.N * BOX_FAC * dm.ns);
    bb.act = alloc<uint8_t>((size_t)dm.NW * dm.ns);
    bb.act32 = alloc<uint32_t>((size_t)dm.NW * dm.ns);
    bb.iters = alloc<int32_t>(dm.ns);
    bb.nactive = alloc<int32_t>(dm.ns);
    bb.V = nullptr;
    if (g_rows > 0) {
      double W[81], lo[9], hi[9];
      const int m = make_row_basis(c, g_rows, g_row_a, g_row_lb, g_row_ub, W, Vrow, lo, hi);
      if (m > 0) {
        bb.V = Vrow;
        bc.enable = 1;
        bc.general = 1;
        bc.nrows = m;
        bc.max_iter = c.v_box_max_iter > 0 ? c.v_box_max_iter : 400;
        bc.mask9 = (1 << m) - 1;
        for (int r = 0; r < 9; ++r) {
          bc.lo9[r] = r < m ? lo[r] : -1e300;
          bc.hi9[r] = r < m ? hi[r] : 1e300;
        }
      }
    }
    StateSizes s = state_sizes(dm);
    b.ekf_q = alloc<T>(s.ekf_q);
    b.ekf_P = alloc<T>(s.ekf_P);
    b.ekf_hist = alloc<T>(s.ekf_hist);
    b.ekf_hist_time = alloc<double>(s.ekf_hist_time);
    b.arr_P = alloc<T>(s.arr_P);
    b.arr_x = alloc<T>(s.arr_x);
    b.win = alloc<T>(s.win);
    b.hist_time = alloc<double>(s.hist_time);
    b.hist_quat = alloc<double>(s.hist_quat);
    b.wp = alloc<double>(s.wp);
    b.wp_time = alloc<double>(s.wp_time);
    b.wp_count = alloc<int32_t>(s.wp_count);
    b.p_vo = alloc<double>(s.p_vo);
    b.pend_flag = alloc<uint8_t>(s.pend_flag);
    b.pend = alloc<double>(s.pend);
    b.status = alloc<int32_t>(s.status);
    fc = make_foot_const(c);
    {
      const int L = robot_num_legs(c.robot), DS = 9 + 3 * L;
      fb.leg = alloc<double>((size_t)dm.NW * (9 * L + 1) * dm.ns);
      fb.arr_M = alloc<double>((size_t)(DS * (DS + 1) / 2) * dm.ns);
      fb.arr_m = alloc<double>((size_t)DS * dm.ns);
      b.foot_leg = c.leg_odom_type == 1 ? fb.leg : nullptr;
    }
    b.ckpt = alloc<T>((size_t)dm.NW * 54 * dm.ns);
    b.resweep = alloc<int32_t>(2 * (size_t)dm.ns);
    const int n = dm.ns;
    for (int i = 0; i < dm.n; ++i) {
      for (int f = 0; f < 4; ++f) {
        b.ekf_q[(size_t)f * n + i] = ec.q0[f];
        b.ekf_P[(size_t)(f * 5) * n + i] = ec.P0[f];
      }
      const int diag[3] = {0, 3, 5};
      for (int f = 0; f < 3; ++f) {
        b.arr_P[(size_t)(0 + diag[f]) * n + i] = mc.P0[f];
        b.arr_P[(size_t)(6 + diag[f]) * n + i] = mc.P0[3 + f];
        b.arr_P[(size_t)(12 + diag[f]) * n + i] = mc.P0[6 + f];
      }
    }
  }
};

template <typename T, typename Model>
void run(const dekf_config &cfg, int S, const double *gyro, const double *accel, const double *imu_time,
         const double *joint_pos, const double *joint_vel, const double *foot_force, const uint8_t *vo_flag,
         const double *vo_quat, const double *vo_time_pre, const double *vo_time_now, const double *vo_rel_p,
         const double *quat_in, double *quat_out, double *x_out, double *vb_out, uint8_t *contact_out,
         int32_t *vo_dbg, int32_t *ekf_dbg, double *pvo_out, int32_t *status_out, double *arrP_out, double *arrx_out,
         int32_t *qp_out) {
  Sim<T> sim(cfg);
  const int n = cfg.n_instances;
  const int nq = Model::NLEG * Model::NJ, nl = Model::NLEG;
  for (int s = 0; s < S; ++s) {
    Inputs in;
    in.gyro = gyro + (size_t)s * 3 * n;
    in.accel = accel + (size_t)s * 3 * n;
    in.imu_time = imu_time + (size_t)s * n;
    in.joint_pos = joint_pos + (size_t)s * nq * n;
    in.joint_vel = joint_vel + (size_t)s * nq * n;
    in.foot_force = foot_force + (size_t)s * nl * n;
    in.vo_flag = vo_flag ? vo_flag + (size_t)s * n : nullptr;
    in.vo_quat = vo_quat + (size_t)s * 4 * n;
    in.vo_time_pre = vo_time_pre + (size_t)s * n;
    in.vo_time_now = vo_time_now + (size_t)s * n;
    in.vo_rel_p = vo_rel_p + (size_t)s * 3 * n;
    in.quat = quat_in ? quat_in + (size_t)s * 4 * n : nullptr;
    Outputs out;
    std::memset(&out, 0, sizeof(out));
    out.quat = quat_out + (size_t)s * 4 * n;
    const int xr = cfg.leg_odom_type == 1 ? 9 + 3 * nl : 9;
    out.x = x_out + (size_t)s * xr * n;
    out.v_body = vb_out + (size_t)s * 3 * n;
    out.contact = contact_out + (size_t)s * nl * n;
    out.dbg_vo = vo_dbg + (size_t)s * 8 * n;
    out.dbg_ekf = ekf_dbg + (size_t)s * 3 * n;
    for (int i = 0; i < n; ++i) {
      int st = ekf_tick<T>(sim.ec, sim.dm, sim.b, in, out, s, i);
      double q[4];
      for (int f = 0; f < 4; ++f)
        q[f] = in.quat ? in.quat[(size_t)f * n + i] : (double)sim.b.ekf_q[(size_t)f * sim.dm.ns + i];
      st |= mhe_assemble<T, Model>(sim.mc, sim.dm, sim.b, in, out, s, i, q);
      if (cfg.leg_odom_type == 1) {
        if (cfg.est_type == 1 || s >= 1) st |= foot_solve<T, Model::NLEG>(sim.fc, sim.dm, sim.b, sim.fb, in, out, s, i);
      } else if (cfg.est_type == 1)
        st |= kf_update<T>(sim.mc, sim.dm, sim.b, in, out, s, i);
      else if (s >= 1 && sim.bc.enable)
        st |= mhe_solve_box<T>(sim.mc, sim.bc, sim.dm, sim.b, sim.bb, in, out, s, i);
      else if (s >= 1 && sim.mc.window_solve == 1)
        st |= mhe_solve_incr<T>(sim.mc, sim.dm, sim.b, in, out, s, i);
      else if (s >= 1)
        st |= mhe_solve<T>(sim.mc, sim.dm, sim.b, in, out, s, i);
      if (qp_out) {
        qp_out[((size_t)s * 2 + 0) * n + i] = sim.bc.enable && s >= 1 ? sim.bb.iters[i] : 0;
        qp_out[((size_t)s * 2 + 1) * n + i] = sim.bc.enable && s >= 1 ? sim.bb.nactive[i] : 0;
      }
      status_out[(size_t)s * n + i] = st;
      for (int f = 0; f < 3; ++f) pvo_out[((size_t)s * 3 + f) * n + i] = sim.b.p_vo[(size_t)f * sim.dm.ns + i];
    }
  }
  if (sim.mc.window_solve == 1)
    for (int i = 0; i < n; ++i) arrival_from_checkpoint<T>(sim.mc, sim.dm, sim.b, S - 1, i);
  for (int f = 0; f < 45; ++f)
    for (int i = 0; i < n; ++i) arrP_out[(size_t)f * n + i] = (double)sim.b.arr_P[(size_t)f * sim.dm.ns + i];
  for (int f = 0; f < 9; ++f)
    for (int i = 0; i < n; ++i) arrx_out[(size_t)f * n + i] = (double)sim.b.arr_x[(size_t)f * sim.dm.ns + i];
}
}  // namespace

extern "C" {
void hostsim_default_go1(dekf_config *c) { fill_go1_defaults(c); }
void hostsim_set_rows(int count, const double *a, const double *lb, const double *ub) {
  g_rows = count > 9 ? 9 : (count < 0 ? 0 : count);
  for (int i = 0; i < g_rows; ++i) {
    for (int k = 0; k < 9; ++k) g_row_a[i * 9 + k] = a[i * 9 + k];
    g_row_lb[i] = lb[i];
    g_row_ub[i] = ub[i];
  }
}
int hostsim_run(const dekf_config *cfg, int S, const double *gyro, const double *accel, const double *imu_time,
                const double *joint_pos, const double *joint_vel, const double *foot_force, const uint8_t *vo_flag,
                const double *vo_quat, const double *vo_time_pre, const double *vo_time_now, const double *vo_rel_p,
                const double *quat_in, double *quat_out, double *x_out, double *vb_out, uint8_t *contact_out,
                int32_t *vo_dbg, int32_t *ekf_dbg, double *pvo_out, int32_t *status_out, double *arrP_out,
                double *arrx_out, int32_t *qp_out) {
#define ARGS *cfg, S, gyro, accel, imu_time, joint_pos, joint_vel, foot_force, vo_flag, vo_quat, vo_time_pre, \
             vo_time_now, vo_rel_p, quat_in, quat_out, x_out, vb_out, contact_out, vo_dbg, ekf_dbg, pvo_out,  \
             status_out, arrP_out, arrx_out, qp_out
  const bool f32 = cfg->precision == DEKF_FP32;
  switch (cfg->robot) {
    case DEKF_ROBOT_GO1:
      if (f32) run<float, Go1Model<float>>(ARGS); else run<double, Go1Model<double>>(ARGS);
      return 0;
    case DEKF_ROBOT_CASSIE:
      if (f32) run<float, CassieModel<float>>(ARGS); else run<double, CassieModel<double>>(ARGS);
      return 0;
    case DEKF_ROBOT_POGOX:
      if (f32) run<float, PogoXModel<float>>(ARGS); else run<double, PogoXModel<double>>(ARGS);
      return 0;
  }
  return -1;
}
}
