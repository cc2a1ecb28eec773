/* CPU oracle: orientation EKF.  TEST INFRASTRUCTURE ONLY.
 * Literal restatement of /root/reference/src/orien_est/src/orien_ekf.cpp (quirks included:
 * the W indexing bug at :285-291 and the replay off-by-one at :186-205). */
#include "oracle.h"
#include "la.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_ekf {
  double dt;
  double gravity[3];
  double Cov_q[16], C_gyro[9], C_accel[9], C_vo[16];
  double q[4];
  /* unbounded history stacks, orien_ekf.hpp:46-53 */
  int discrete_time;
  int n, cap;
  double *gyro_stack, *accel_stack, *time_stack, *q_stack, *P_stack;
  int *disc_stack;
  int last_cur, last_idx, last_nreplay;
};

/* orien_ekf.cpp:8-33 */
orc_ekf *orc_ekf_create(const double init_std[4], const double process_std[3],
                        const double gravity_meas_std[3], const double vo_meas_std[4],
                        const double quaternion_init[4], int rate) {
  orc_ekf *e = (orc_ekf *)calloc(1, sizeof(orc_ekf));
  e->gravity[0] = 0;
  e->gravity[1] = 0;
  e->gravity[2] = 9.81;                 /* :11 */
  e->dt = 1 / (double)rate;             /* :25 */
  for (int i = 0; i < 4; ++i) e->Cov_q[i * 4 + i] = pow(init_std[i], 2);     /* :27 */
  for (int i = 0; i < 3; ++i) e->C_gyro[i * 3 + i] = pow(process_std[i], 2); /* :28 */
  for (int i = 0; i < 3; ++i) e->C_accel[i * 3 + i] = pow(gravity_meas_std[i], 2); /* :29 */
  for (int i = 0; i < 4; ++i) e->C_vo[i * 4 + i] = pow(vo_meas_std[i], 2);   /* :30 */
  for (int i = 0; i < 4; ++i) e->q[i] = quaternion_init[i];                   /* :31-33 */
  e->last_cur = e->last_idx = -1;
  return e;
}

void orc_ekf_destroy(orc_ekf *e) {
  if (!e) return;
  free(e->gyro_stack);
  free(e->accel_stack);
  free(e->time_stack);
  free(e->q_stack);
  free(e->P_stack);
  free(e->disc_stack);
  free(e);
}

void orc_ekf_get(const orc_ekf *e, double q[4], double P[16]) {
  if (q) memcpy(q, e->q, sizeof(double) * 4);
  if (P) memcpy(P, e->Cov_q, sizeof(double) * 16);
}
void orc_ekf_set(orc_ekf *e, const double q[4], const double P[16]) {
  if (q) memcpy(e->q, q, sizeof(double) * 4);
  if (P) memcpy(e->Cov_q, P, sizeof(double) * 16);
}
void orc_ekf_last_replay(const orc_ekf *e, int *cur, int *idx, int *nreplay) {
  *cur = e->last_cur;
  *idx = e->last_idx;
  *nreplay = e->last_nreplay;
}

/* orien_ekf.cpp:214-228 */
static void gyro_2_Ohm(const double w[3], double Ohm[16]) {
  la_zero(Ohm, 16);
  Ohm[0 * 4 + 1] = -w[0];
  Ohm[0 * 4 + 2] = -w[1];
  Ohm[0 * 4 + 3] = -w[2];
  Ohm[1 * 4 + 0] = w[0];
  Ohm[2 * 4 + 0] = w[1];
  Ohm[3 * 4 + 0] = w[2];
  Ohm[1 * 4 + 2] = w[2];
  Ohm[1 * 4 + 3] = -w[1];
  Ohm[2 * 4 + 3] = w[0];
  Ohm[2 * 4 + 1] = -w[2];
  Ohm[3 * 4 + 1] = w[1];
  Ohm[3 * 4 + 2] = -w[0];
}

/* orien_ekf.cpp:270-294 -- including the indexing bug: the last two writes go to row 2 again,
 * so row 2 ends as (z, x, w) and row 3 as (-y, 0, 0). */
static void quat_2_W(const double q[4], double dt, double W[12]) {
  la_zero(W, 12);
  W[0 * 3 + 0] = -q[1];
  W[0 * 3 + 1] = -q[2];
  W[0 * 3 + 2] = -q[3];

  W[1 * 3 + 0] = q[0];
  W[1 * 3 + 1] = -q[3];
  W[1 * 3 + 2] = q[2];

  W[2 * 3 + 0] = q[3];
  W[2 * 3 + 1] = q[0];
  W[2 * 3 + 2] = -q[1];

  W[3 * 3 + 0] = -q[2];
  W[2 * 3 + 1] = q[1]; /* sic, :290 */
  W[2 * 3 + 2] = q[0]; /* sic, :291 */
  for (int i = 0; i < 12; ++i) W[i] = 0.5 * dt * W[i];
}

/* Eigen::Quaterniond::normalized().toRotationMatrix() (orien_ekf.cpp:296-305, DecentralEst.cpp:867) */
void orc_quat_to_rot(const double qin[4], double R[9]) {
  double nrm = sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
  double w = qin[0] / nrm, x = qin[1] / nrm, y = qin[2] / nrm, z = qin[3] / nrm;
  double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  double twx = tx * w, twy = ty * w, twz = tz * w;
  double txx = tx * x, txy = ty * x, txz = tz * x;
  double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz);
  R[1] = txy - twz;
  R[2] = txz + twy;
  R[3] = txy + twz;
  R[4] = 1 - (txx + tzz);
  R[5] = tyz - twx;
  R[6] = txz - twy;
  R[7] = tyz + twx;
  R[8] = 1 - (txx + tyy);
}

/* orien_ekf.cpp:307-329 */
static void quat_2_H(const double q[4], const double g[3], double H[12]) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  H[0 * 4 + 0] = g[0] * w + g[1] * z - g[2] * y;
  H[0 * 4 + 1] = g[0] * x + g[1] * y + g[2] * z;
  H[0 * 4 + 2] = -g[0] * y + g[1] * x - g[2] * w;
  H[0 * 4 + 3] = -g[0] * z + g[1] * w + g[2] * x;

  H[1 * 4 + 0] = -g[0] * z + g[1] * w + g[2] * x;
  H[1 * 4 + 1] = g[0] * y - g[1] * x + g[2] * w;
  H[1 * 4 + 2] = g[0] * x + g[1] * y + g[2] * z;
  H[1 * 4 + 3] = -g[0] * w - g[1] * z + g[2] * y;

  H[2 * 4 + 0] = g[0] * y - g[1] * x + g[2] * w;
  H[2 * 4 + 1] = g[0] * z - g[1] * w - g[2] * x;
  H[2 * 4 + 2] = g[0] * w + g[1] * z - g[2] * y;
  H[2 * 4 + 3] = g[0] * x + g[1] * y + g[2] * z;
  for (int i = 0; i < 12; ++i) H[i] = 2 * H[i];
}

/* orien_ekf.cpp:353-357 */
static void quat_norm(double q[4]) {
  double n = la_norm2(q, 4);
  for (int i = 0; i < 4; ++i) q[i] = q[i] / n;
}

/* orien_ekf.cpp:108-123 */
void orc_ekf_predict(const orc_ekf *e, double q_pred[4], const double q[4], const double gyro[3],
                     const double P[16], double P_pred[16]) {
  double Ohm[16], W[12], F[16], FP[16], FPFt[16], WC[12], WCWt[16];
  gyro_2_Ohm(gyro, Ohm);
  quat_2_W(q, e->dt, W);
  for (int i = 0; i < 16; ++i) F[i] = ((i % 5 == 0) ? 1.0 : 0.0) + e->dt / 2 * Ohm[i];
  double qp[4];
  la_mm(qp, F, q, 4, 4, 1);
  la_mm(FP, F, P, 4, 4, 4);
  la_mmt(FPFt, FP, F, 4, 4, 4);
  la_mm(WC, W, e->C_gyro, 4, 3, 3);
  la_mmt(WCWt, WC, W, 4, 3, 4);
  for (int i = 0; i < 16; ++i) P_pred[i] = FPFt[i] + WCWt[i];
  quat_norm(qp);
  memcpy(q_pred, qp, sizeof(qp));
}

/* orien_ekf.cpp:125-142 */
void orc_ekf_correct(const orc_ekf *e, double q_corr[4], const double q_pred[4],
                     const double accel[3], const double P_pred[16], double P_corr[16]) {
  double R[9], Rt[9], a_hat[3], H[12], PHt[12], S[9], Sinv[9], K[12], KH[16], IKH[16];
  orc_quat_to_rot(q_pred, R);
  la_transpose(Rt, R, 3, 3);
  la_mm(a_hat, Rt, e->gravity, 3, 3, 1);
  quat_2_H(q_pred, e->gravity, H);
  double rel = la_norm2(accel, 3) / la_norm2(e->gravity, 3);
  la_mmt(PHt, P_pred, H, 4, 4, 3); /* P H^T : 4x3 */
  la_mm(S, H, PHt, 3, 4, 3);
  for (int i = 0; i < 9; ++i) S[i] += rel * rel * e->C_accel[i];
  la_inverse(Sinv, S, 3);
  la_mm(K, PHt, Sinv, 4, 3, 3);
  double innov[3] = {accel[0] - a_hat[0], accel[1] - a_hat[1], accel[2] - a_hat[2]};
  double dq[4];
  la_mm(dq, K, innov, 4, 3, 1);
  double qc[4];
  for (int i = 0; i < 4; ++i) qc[i] = q_pred[i] + dq[i];
  la_mm(KH, K, H, 4, 3, 4);
  for (int i = 0; i < 16; ++i) IKH[i] = ((i % 5 == 0) ? 1.0 : 0.0) - KH[i];
  double Pc[16];
  la_mm(Pc, IKH, P_pred, 4, 4, 4);
  memcpy(P_corr, Pc, sizeof(Pc));
  quat_norm(qc);
  memcpy(q_corr, qc, sizeof(qc));
}

/* orien_ekf.cpp:144-154 (H = I4) */
void orc_ekf_vo_correct(const orc_ekf *e, double q_corr[4], const double q_pred[4],
                        const double q_vo[4], const double P_pred[16], double P_corr[16]) {
  double S[16], Sinv[16], K[16], IK[16];
  for (int i = 0; i < 16; ++i) S[i] = P_pred[i] + e->C_vo[i];
  la_inverse(Sinv, S, 4);
  la_mm(K, P_pred, Sinv, 4, 4, 4);
  double innov[4], dq[4], qc[4];
  for (int i = 0; i < 4; ++i) innov[i] = q_vo[i] - q_pred[i];
  la_mm(dq, K, innov, 4, 4, 1);
  for (int i = 0; i < 4; ++i) qc[i] = q_pred[i] + dq[i];
  for (int i = 0; i < 16; ++i) IK[i] = ((i % 5 == 0) ? 1.0 : 0.0) - K[i];
  double Pc[16];
  la_mm(Pc, IK, P_pred, 4, 4, 4);
  memcpy(P_corr, Pc, sizeof(Pc));
  quat_norm(qc);
  memcpy(q_corr, qc, sizeof(qc));
}

static void push_history(orc_ekf *e, const double gyro[3], const double accel[3], double t) {
  if (e->n == e->cap) {
    e->cap = e->cap ? 2 * e->cap : 1024;
    e->gyro_stack = (double *)realloc(e->gyro_stack, sizeof(double) * 3 * (size_t)e->cap);
    e->accel_stack = (double *)realloc(e->accel_stack, sizeof(double) * 3 * (size_t)e->cap);
    e->time_stack = (double *)realloc(e->time_stack, sizeof(double) * (size_t)e->cap);
    e->q_stack = (double *)realloc(e->q_stack, sizeof(double) * 4 * (size_t)e->cap);
    e->P_stack = (double *)realloc(e->P_stack, sizeof(double) * 16 * (size_t)e->cap);
    e->disc_stack = (int *)realloc(e->disc_stack, sizeof(int) * (size_t)e->cap);
  }
  memcpy(e->gyro_stack + 3 * e->n, gyro, sizeof(double) * 3);
  memcpy(e->accel_stack + 3 * e->n, accel, sizeof(double) * 3);
  e->time_stack[e->n] = t;
  e->disc_stack[e->n] = e->discrete_time;
  memcpy(e->q_stack + 4 * e->n, e->q, sizeof(double) * 4);
  memcpy(e->P_stack + 16 * e->n, e->Cov_q, sizeof(double) * 16);
  e->n++;
}

/* std::upper_bound: index of the first element > v */
static int upper_bound_d(const double *a, int n, double v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = lo + (hi - lo) / 2;
    if (!(v < a[mid]))
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

/* orien_ekf.cpp:156-212 */
static void get_measurement(orc_ekf *e, const double gyro[3], const double accel[3], double imu_time,
                            int vo_new, const double vo_quat[4], double vo_time) {
  push_history(e, gyro, accel, imu_time); /* :158-163 */
  if (vo_new && e->n > 0) {               /* :165 */
    int ub = upper_bound_d(e->time_stack, e->n, vo_time); /* :175 */
    e->last_cur = e->n - 1;
    e->last_nreplay = 0;
    if (ub == 0) {
      e->last_idx = -1; /* :178-183 drop */
    } else {
      int idx = ub - 1;                                                   /* :186 */
      int rel = e->disc_stack[e->n - 1] - e->disc_stack[idx];             /* :187 */
      e->last_idx = idx;
      memcpy(e->q, e->q_stack + 4 * idx, sizeof(double) * 4);             /* :188 */
      memcpy(e->Cov_q, e->P_stack + 16 * idx, sizeof(double) * 16);       /* :189 */
      double q_pred[4], q_corr[4], P_pred[16], P_corr[16];
      for (int i = 0; i < rel - 1; ++i) {                                 /* :191 */
        orc_ekf_predict(e, q_pred, e->q, e->gyro_stack + 3 * (idx + i), e->Cov_q, P_pred);
        orc_ekf_correct(e, q_corr, q_pred, e->accel_stack + 3 * (idx + i), P_pred, P_corr);
        if (i == 0) {                                                     /* :197-202 */
          memcpy(q_pred, q_corr, sizeof(q_pred));
          memcpy(P_pred, P_corr, sizeof(P_pred));
          orc_ekf_vo_correct(e, q_corr, q_pred, vo_quat, P_pred, P_corr);
        }
        memcpy(e->q, q_corr, sizeof(q_corr));
        memcpy(e->Cov_q, P_corr, sizeof(P_corr));
        e->last_nreplay++;
      }
    }
  }
}

/* orien_ekf.cpp:77-89 (init_imu is true: the caller supplies an IMU sample every tick) */
void orc_ekf_tick(orc_ekf *e, const double gyro[3], const double accel[3], double imu_time,
                  int vo_new, const double vo_quat[4], double vo_time) {
  double q_pred[4], q_corr[4], P_pred[16], P_corr[16];
  get_measurement(e, gyro, accel, imu_time, vo_new, vo_quat, vo_time);
  orc_ekf_predict(e, q_pred, e->q, gyro, e->Cov_q, P_pred);
  orc_ekf_correct(e, q_corr, q_pred, accel, P_pred, P_corr);
  memcpy(e->q, q_corr, sizeof(q_corr));
  memcpy(e->Cov_q, P_corr, sizeof(P_corr));
  e->discrete_time++;
}
