/* CPU oracle: batch driver.  TEST INFRASTRUCTURE ONLY.
 * Steps many independent (EKF + MHE) oracle instances over SoA sensor streams laid out exactly as
 * the device ABI takes them ([step][field][instance]), one instance per thread at a time
 * (BASELINE.md section 2 protocol).  Call order per tick follows the reference's two timers run in
 * lock-step: orien_ekf::timerCallback (orien_ekf.cpp:77-89) then robotSub::timerCallback
 * (EstSub.cpp:58-75) consuming the freshly published quaternion. */
#define _GNU_SOURCE
#include "oracle.h"
#include <pthread.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct {
  int n, S, nq, nlegs;
  const double *gyro, *accel, *imu_time, *joint_pos, *joint_vel, *foot_force;
  const unsigned char *vo_flag;
  const double *vo_quat, *vo_time_pre, *vo_time_now, *vo_rel_p;
  const double *quat_in; /* optional external orientation for the MHE, [S][4][n] */
} orc_stream;

typedef struct {
  double init_std[4], process_std[3], gravity_meas_std[3], vo_meas_std[4], quaternion_init[4];
  int rate;
} orc_ekf_params;

typedef struct {
  /* all optional (NULL to skip); [S][k][n] */
  double *quat, *x, *v_body, *p_vo;
  unsigned char *contact;
  int *vo_dbg; /* [S][10][n] */
  int *ekf_dbg; /* [S][3][n]: cur, idx, nreplay of the replay processed at that tick (or -2) */
  double *M_p, *n_p; /* final arrival cost only: [ds*ds][n], [ds][n] */
  int *admm_iters;   /* [S][n] */
} orc_outputs;

typedef struct {
  const orc_params *prm;
  const orc_ekf_params *eprm;
  const orc_stream *st;
  orc_outputs *out;
  int i0, i1, run_ekf, run_mhe;
  int t_steady; /* first step counted in the timing */
  int tid;
  double seconds;
} job_t;

/* bench.py's CPU baseline pins worker t to the t-th CPU of the process's affinity mask (stable timings on a busy box) */
static int g_pin_threads = 0;
void orc_set_pin_threads(int on) { g_pin_threads = on; }
static void pin_to_allowed_cpu(int t) {
  cpu_set_t allowed, one;
  if (sched_getaffinity(0, sizeof(allowed), &allowed) != 0) return;
  int cnt = CPU_COUNT(&allowed);
  if (cnt <= 0) return;
  int want = t % cnt, seen = 0;
  for (int c = 0; c < CPU_SETSIZE; ++c)
    if (CPU_ISSET(c, &allowed)) {
      if (seen == want) {
        CPU_ZERO(&one);
        CPU_SET(c, &one);
        pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
        return;
      }
      ++seen;
    }
}

static double now_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void *worker(void *arg) {
  job_t *jb = (job_t *)arg;
  if (g_pin_threads) pin_to_allowed_cpu(jb->tid);
  const orc_stream *st = jb->st;
  int n = st->n, S = st->S, nq = st->nq, nl = st->nlegs;
  int ds, dm, dc, nV, nC;
  double t_acc = 0.0;
  for (int i = jb->i0; i < jb->i1; ++i) {
    orc_ekf *e = NULL;
    orc_mhe *m = NULL;
    if (jb->run_ekf)
      e = orc_ekf_create(jb->eprm->init_std, jb->eprm->process_std, jb->eprm->gravity_meas_std,
                         jb->eprm->vo_meas_std, jb->eprm->quaternion_init, jb->eprm->rate);
    if (jb->run_mhe) {
      m = orc_mhe_create(jb->prm);
      orc_mhe_get_dims(m, &ds, &dm, &dc, &nV, &nC);
    }
    double t0 = 0.0;
    for (int s = 0; s < S; ++s) {
      if (s == jb->t_steady) t0 = now_s();
      size_t b1 = (size_t)s * n + i;
      double gyro[3], accel[3], q[4] = {1, 0, 0, 0}, vq[4] = {1, 0, 0, 0};
      for (int c = 0; c < 3; ++c) {
        gyro[c] = st->gyro[((size_t)s * 3 + c) * n + i];
        accel[c] = st->accel[((size_t)s * 3 + c) * n + i];
      }
      int vo_new = st->vo_flag ? st->vo_flag[b1] : 0;
      double t_imu = st->imu_time[b1];
      if (e) {
        if (vo_new)
          for (int c = 0; c < 4; ++c) vq[c] = st->vo_quat[((size_t)s * 4 + c) * n + i];
        orc_ekf_tick(e, gyro, accel, t_imu, vo_new, vq, vo_new ? st->vo_time_now[b1] : 0.0);
        orc_ekf_get(e, q, NULL);
        if (jb->out->quat)
          for (int c = 0; c < 4; ++c) jb->out->quat[((size_t)s * 4 + c) * n + i] = q[c];
        if (jb->out->ekf_dbg) {
          int cur = -2, idx = -2, nr = -2;
          if (vo_new) orc_ekf_last_replay(e, &cur, &idx, &nr);
          jb->out->ekf_dbg[((size_t)s * 3 + 0) * n + i] = cur;
          jb->out->ekf_dbg[((size_t)s * 3 + 1) * n + i] = idx;
          jb->out->ekf_dbg[((size_t)s * 3 + 2) * n + i] = nr;
        }
      }
      if (m) {
        orc_sample smp;
        memset(&smp, 0, sizeof(smp));
        smp.imu_time = t_imu;
        for (int c = 0; c < 3; ++c) {
          smp.accel_b[c] = accel[c];
          smp.angular_b[c] = gyro[c];
        }
        if (st->quat_in)
          for (int c = 0; c < 4; ++c) smp.quaternion[c] = st->quat_in[((size_t)s * 4 + c) * n + i];
        else
          for (int c = 0; c < 4; ++c) smp.quaternion[c] = q[c];
        for (int c = 0; c < nq; ++c) {
          smp.joint_pos[c] = st->joint_pos[((size_t)s * nq + c) * n + i];
          smp.joint_vel[c] = st->joint_vel[((size_t)s * nq + c) * n + i];
        }
        for (int c = 0; c < nl; ++c) smp.joint_pos[nq + c] = st->foot_force[((size_t)s * nl + c) * n + i];
        smp.vo_new = vo_new;
        if (vo_new) {
          smp.vo_time_pre = st->vo_time_pre[b1];
          smp.vo_time_now = st->vo_time_now[b1];
          for (int c = 0; c < 3; ++c) smp.vo_p[c] = st->vo_rel_p[((size_t)s * 3 + c) * n + i];
        }
        orc_mhe_step(m, s, &smp);
        if (jb->out->x && s >= 1) {
          double x[32];
          if (jb->prm->est_type == 0)
            orc_mhe_get_x(m, x);
          else {
            double C[32 * 32], vb[3];
            orc_mhe_get_kf(m, x, C, vb);
          }
          for (int c = 0; c < ds; ++c) jb->out->x[((size_t)s * ds + c) * n + i] = x[c];
        }
        if (jb->out->v_body && s >= 1) {
          double v[3], x[32], C[32 * 32];
          if (jb->prm->est_type == 0)
            orc_mhe_get_v_body(m, v);
          else
            orc_mhe_get_kf(m, x, C, v);
          for (int c = 0; c < 3; ++c) jb->out->v_body[((size_t)s * 3 + c) * n + i] = v[c];
        }
        if (jb->out->p_vo) {
          double p[3];
          orc_mhe_get_p_vo(m, p);
          for (int c = 0; c < 3; ++c) jb->out->p_vo[((size_t)s * 3 + c) * n + i] = p[c];
        }
        if (jb->out->contact) {
          double ct[8];
          orc_mhe_get_contact(m, ct);
          for (int c = 0; c < nl; ++c) jb->out->contact[((size_t)s * nl + c) * n + i] = (unsigned char)(ct[c] != 0.0);
        }
        if (jb->out->vo_dbg) {
          int dbg[10];
          orc_mhe_get_vo_debug(m, dbg);
          for (int c = 0; c < 10; ++c) jb->out->vo_dbg[((size_t)s * 10 + c) * n + i] = vo_new ? dbg[c] : -2;
        }
        if (jb->out->admm_iters) jb->out->admm_iters[b1] = orc_mhe_get_admm_iters(m);
      }
    }
    if (jb->t_steady < S) t_acc += now_s() - t0;
    if (m && jb->out->M_p) {
      double M[32 * 32], nn[32];
      if (orc_mhe_get_arrival(m, M, nn)) {
        for (int c = 0; c < ds * ds; ++c) jb->out->M_p[(size_t)c * n + i] = M[c];
        for (int c = 0; c < ds; ++c) jb->out->n_p[(size_t)c * n + i] = nn[c];
      }
    }
    if (e) orc_ekf_destroy(e);
    if (m) orc_mhe_destroy(m);
  }
  jb->seconds = t_acc;
  return NULL;
}

/* Runs instances [i0, i1) with nthreads threads.  Returns wall-clock seconds of the whole call;
 * *busy_seconds (optional) receives the per-thread time spent in steps >= t_steady summed over
 * threads (so steps/s/core = (i1-i0)*(S-t_steady)/busy_seconds). */
double orc_run_batch(const orc_params *prm, const orc_ekf_params *eprm, const orc_stream *st,
                     orc_outputs *out, int i0, int i1, int nthreads, int run_ekf, int run_mhe,
                     int t_steady, double *busy_seconds, double *busy_max) {
  if (nthreads < 1) nthreads = 1;
  int cnt = i1 - i0;
  if (nthreads > cnt) nthreads = cnt > 0 ? cnt : 1;
  job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t *th = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  double t0 = now_s();
  for (int t = 0; t < nthreads; ++t) {
    jobs[t].prm = prm;
    jobs[t].eprm = eprm;
    jobs[t].st = st;
    jobs[t].out = out;
    jobs[t].i0 = i0 + (int)((long long)cnt * t / nthreads);
    jobs[t].i1 = i0 + (int)((long long)cnt * (t + 1) / nthreads);
    jobs[t].run_ekf = run_ekf;
    jobs[t].run_mhe = run_mhe;
    jobs[t].t_steady = t_steady;
    jobs[t].tid = t;
    pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  double busy = 0.0, bmax = 0.0;
  for (int t = 0; t < nthreads; ++t) {
    pthread_join(th[t], NULL);
    busy += jobs[t].seconds;
    if (jobs[t].seconds > bmax) bmax = jobs[t].seconds;
  }
  if (busy_max) *busy_max = bmax;
  double wall = now_s() - t0;
  if (busy_seconds) *busy_seconds = busy;
  free(jobs);
  free(th);
  return wall;
}
