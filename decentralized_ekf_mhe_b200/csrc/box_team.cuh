// State-constrained window solve, team form: 9 lanes of a warp per estimator instance (3 instances per warp).
//
// Same algorithm and same operand values as box_solve() in box_solve.cuh (primal-dual active set over a
// block-tridiagonal Cholesky recursion with 9x9 blocks; reference mechanism: MHEproblem::addConstraints(name, lb, ub),
// decentral_legged_est/src/MheSrb.cpp:58-68) -- re-laid out for a batch that is too small for one thread per instance
// (BASELINE config 4: 16,384 instances = 4 warps per SM, every thread a serial chain through 6 KB of local arrays).
// Lane r of a team owns ROW r of every 9x9 block in registers; rows of other lanes arrive by warp shuffle:
//   * stage assembly: every lane builds its own row of A'QA, E, E' and Q in closed form from the stage record
//     (row r of Q and of A'Q by row type p / v / b, then `row x A` with A's block structure);
//   * Schur update  D_j -= F_{j-1} F_{j-1}'  : 81 row-element broadcasts;
//   * Cholesky, the triangular solve F_j = E_j' L_j^-T and the forward substitution y_j = L_j^-1 rhs_j run in ONE
//     right-looking loop over the 9 columns: column j of L is broadcast once and used by all three;
//   * factors (L_j, F_j, y_j) go to a per-instance contiguous scratch [instance][stage][135]; the backward substitution
//     reads COLUMNS of L_j and F_j from it (lane c owns x_j[c]);
//   * multipliers / active-set update: every lane of the team evaluates the same scalar formulas (no divergence), lane 0
//     writes the masks.
// All lanes of a warp execute the same instruction stream (full-mask shuffles); teams that have converged keep
// re-solving with their final active set (idempotent) until the slowest team of the warp is done.
#pragma once
#include "box_solve.cuh"

namespace dekf {

enum { BOX_PRIOR = 90 };  // per instance: symmetrised M = Pa^-1 (81, row-major) + m = M xa (9)
// per instance and stage: L (45, packed lower, inverse diagonal) + F (81) + y (9) + x (9) + the v rows of the UNMODIFIED
// blocks for the multipliers: 3 x [D0 row (9), E row (9), E' row (9), r0 (1)] + the carry INTO the stage for a restarted
// forward pass: 9 x [Dc row (9), rc, rc0]
enum { BOX_TFAC_Y = 126, BOX_TFAC_X = 135, BOX_TFAC_GRAD = 144, BOX_TFAC_CARRY = 144 + 3 * 28, BOX_TFAC = 144 + 3 * 28 + 9 * 11 + 1 };

struct BoxTeamBuffers {
  double *prior;  // [ns][90]
  double *fac;    // [ns][N][BOX_TFAC]  (instance-major; replaces BoxBuffers::fac)
};

#if defined(__CUDACC__)

__device__ __forceinline__ double bt_shfl(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
// the three passes are chains of dependent stages with one global-memory round trip at the head of each stage: the data
// of the NEXT stage is requested while the current one computes (no registers held)
__device__ __forceinline__ void bt_prefetch(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// asynchronous global -> shared copy of one element (4 or 8 bytes): the stage record of the NEXT stage is staged while the
// current stage computes, without holding registers (LDGSTS)
template <typename T>
__device__ __forceinline__ void bt_cp_async(T *smem_dst, const T *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc));
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void bt_cp_commit() { asm volatile("cp.async.commit_group;"); }
__device__ __forceinline__ void bt_cp_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <typename T>
__device__ __forceinline__ void bt_stage_record(T *dst /*[REC_SIZE]*/, const T *rec, size_t ns, int r) {
  bt_cp_async(dst + r, rec + (size_t)r * ns);
  bt_cp_async(dst + r + 9, rec + (size_t)(r + 9) * ns);
  if (r + 18 < REC_SIZE) bt_cp_async(dst + r + 18, rec + (size_t)(r + 18) * ns);
  bt_cp_commit();
}
__device__ __forceinline__ double bt_sel3(int m, double a0, double a1, double a2) { return m == 0 ? a0 : (m == 1 ? a1 : a2); }

// v (row vector, 9) times A = [[I, dt I, -h R],[0, I, -dt R],[0, 0, I]]
__device__ __forceinline__ void bt_row_times_A(const double *v, const double *R, double dt, double h, double *out) {
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    out[c] = v[c];
    out[3 + c] = dt * v[c] + v[3 + c];
  }
  double w[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) w[k] = -h * v[k] - dt * v[3 + k];
#pragma unroll
  for (int c = 0; c < 3; ++c) out[6 + c] = w[0] * R[0 * 3 + c] + w[1] * R[1 * 3 + c] + w[2] * R[2 * 3 + c] + v[6 + c];
}

// Column JJ of the right-looking Cholesky of S (row r in D), fused with y = L^-1 rhs and F = G L^-T; recursion instead of
// a loop so that every index into the register rows is a compile-time constant.
template <int JJ>
struct BtFactor {
  static __device__ __forceinline__ void run(double (&D)[9], double (&G)[9], double (&F)[9], double &rhs, double &ymine, int &status,
                                             int base, int r) {
    const double inv = bt_shfl(rsqrt(D[JJ]), base + JJ);  // 1 / L[JJ][JJ]
    if (!(inv > 0.0) || !(inv - inv == 0.0)) status |= ST_NONFINITE;
    double Lrj = D[JJ] * inv;
    if (r == JJ) Lrj = inv;  // the scratch keeps the INVERSE diagonal (the backward substitution multiplies)
    D[JJ] = Lrj;
    const double yj = bt_shfl(rhs * inv, base + JJ);
    if (r == JJ) ymine = yj;
    if (r > JJ) rhs -= Lrj * yj;
    F[JJ] = G[JJ] * inv;
#pragma unroll
    for (int c = JJ + 1; c < 9; ++c) {
      const double Lcj = bt_shfl(Lrj, base + c);
      if (r >= c) D[c] -= Lrj * Lcj;
      G[c] -= F[JJ] * Lcj;
    }
    BtFactor<JJ + 1>::run(D, G, F, rhs, ymine, status, base, r);
  }
};
template <>
struct BtFactor<9> {
  static __device__ __forceinline__ void run(double (&)[9], double (&)[9], double (&)[9], double &, double &, int &, int, int) {}
};

// Active-set solve of one instance by one team.  base: first lane of the team in the warp, r: lane in the team (0..8),
// valid: the team owns a real instance (stores are predicated on it), i: instance index (clamped for idle teams).
// Returns status bits (identical in every lane of the team); x_T[r] in xT_r.
template <typename T>
__device__ int box_team_solve(const BoxConst &bc, const Dims &dm, const Buffers<T> &b, const BoxBuffers &bb, const BoxTeamBuffers &tb,
                              int k0, int Tk, int i, bool valid, int base, int r, T *srec /*shared, [2][REC_SIZE] of this team*/,
                              double &xT_r) {
  const size_t ns = (size_t)dm.ns;
  const int K = Tk - k0 + 1;
  const int slot0 = k0 % dm.NW;
  const int t = r / 3, m = r - 3 * t;
  const double dt = bc.dt, h = 0.5 * bc.dt * bc.dt;
  const double *pri = tb.prior + (size_t)i * BOX_PRIOR;
  double *fac = tb.fac + (size_t)i * dm.N * BOX_TFAC;
  int status = 0;
  int iters = 0, nact = 0;
  bool done = false;
  int first_changed = 0;  // first stage whose active set the previous iteration changed
  xT_r = 0.0;
  const double qa_m = bt_sel3(m, bc.qa[0], bc.qa[1], bc.qa[2]), qb_m = bt_sel3(m, bc.qb[0], bc.qb[1], bc.qb[2]);
  const double qc_m = bt_sel3(m, bc.qc[0], bc.qc[1], bc.qc[2]), qab_m = bt_sel3(m, bc.qab[0], bc.qab[1], bc.qab[2]);

  for (int it = 0; it < bc.max_iter; ++it) {
    if (__all_sync(0xffffffffu, done || !valid)) break;
    if (!done) iters = it + 1;
    // ------------------------------------------------------------------ forward: assemble, active set, factor
    // Stage j's factor depends on the active sets of stages <= j + 1 only: an iteration after the first restarts the
    // forward pass at the stage before the first changed set (warp-uniform: the earliest restart of the warp's teams; a
    // team that restarts earlier than it needs to recomputes what it already has).
    const int j_start = __reduce_min_sync(0xffffffffu, (done || !valid) ? K : (first_changed > 0 ? first_changed - 1 : 0));
    double Dc[9], rc, rc0, Fp[9], yp = 0.0;  // rc0: carry of the right-hand side without the active-set terms
    if (j_start == 0) {
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        Dc[c] = pri[r * 9 + c];
        Fp[c] = 0.0;
      }
      rc = rc0 = pri[81 + r];
    } else {
      const double *cj = fac + (size_t)j_start * BOX_TFAC + BOX_TFAC_CARRY + r * 11;
      const double *fp = fac + (size_t)(j_start - 1) * BOX_TFAC;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        Dc[c] = cj[c];
        Fp[c] = fp[45 + r * 9 + c];
      }
      rc = cj[9];
      rc0 = cj[10];
      yp = fp[BOX_TFAC_Y + r];
    }
    int slot = (slot0 + j_start) % dm.NW;
    // the stage record of the NEXT stage travels global -> shared (cp.async) while the current stage computes
    __syncwarp();
    bt_stage_record(srec, b.win + (size_t)slot * REC_SIZE * ns + i, ns, r);
    int mask_carry = bb.act[(size_t)slot * ns + i];
    for (int j = j_start; j < K; ++j) {
      const int slot_n = slot + 1 == dm.NW ? 0 : slot + 1;
      const bool last = j + 1 >= K;
      const T *rec = srec + ((j - j_start) & 1) * REC_SIZE;  // this stage's record, staged in shared memory
      bt_cp_wait_all();
      __syncwarp();
      if (!last) bt_stage_record(srec + ((j - j_start + 1) & 1) * REC_SIZE, b.win + (size_t)slot_n * REC_SIZE * ns + i, ns, r);
      double R[9], as[3], dlt[3];
#pragma unroll
      for (int f = 0; f < 9; ++f) R[f] = (double)rec[REC_R + f];
#pragma unroll
      for (int f = 0; f < 3; ++f) {
        as[f] = (double)rec[REC_AS + f];
        dlt[f] = (double)rec[REC_DLT + f];
      }
      const bool vo = rec[REC_FLAG] != T(0);
      const int mask = mask_carry;
      if (valid && j > 0) {  // carry into this stage, for a restart of the forward pass here
        double *cj = fac + (size_t)j * BOX_TFAC + BOX_TFAC_CARRY + r * 11;
#pragma unroll
        for (int c = 0; c < 9; ++c) cj[c] = Dc[c];
        cj[9] = rc;
        cj[10] = rc0;
      }
      const int mask_n = last ? 0 : bb.act[(size_t)slot_n * ns + i];
      mask_carry = mask_n;
      double D[9], rhs = rc, rhs0 = rc0, G[9], Dn[9], rn = 0.0, rn0 = 0.0, Erow[9];
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        D[c] = Dc[c];
        G[c] = 0.0;
        Dn[c] = 0.0;
        Erow[c] = 0.0;
      }
      // leg-odometry rows 1/2 v' Lam v - eta' v touch the v rows
      if (t == 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double l0 = (double)rec[REC_LAM + S3<double>::idx(0, c)];
          const double l1 = (double)rec[REC_LAM + S3<double>::idx(1, c)];
          const double l2 = (double)rec[REC_LAM + S3<double>::idx(2, c)];
          D[3 + c] += bt_sel3(m, l0, l1, l2);
        }
        const double eta_m = bt_sel3(m, (double)rec[REC_ETA + 0], (double)rec[REC_ETA + 1], (double)rec[REC_ETA + 2]);
        rhs += eta_m;
        rhs0 += eta_m;
      }
      if (!last) {
        // rows m of S_x = R diag(q_x) R'
        double Rm[3], Rcm[3], Sa[3], Sb[3], Sc[3], Sv[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Rm[c] = bt_sel3(m, R[0 * 3 + c], R[1 * 3 + c], R[2 * 3 + c]);   // R[m][c]
          Rcm[c] = bt_sel3(m, R[c * 3 + 0], R[c * 3 + 1], R[c * 3 + 2]);  // R[c][m]
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Sa[c] = Rm[0] * bc.qa[0] * R[c * 3 + 0] + Rm[1] * bc.qa[1] * R[c * 3 + 1] + Rm[2] * bc.qa[2] * R[c * 3 + 2];
          Sb[c] = Rm[0] * bc.qb[0] * R[c * 3 + 0] + Rm[1] * bc.qb[1] * R[c * 3 + 1] + Rm[2] * bc.qb[2] * R[c * 3 + 2];
          Sc[c] = Rm[0] * bc.qc[0] * R[c * 3 + 0] + Rm[1] * bc.qc[1] * R[c * 3 + 1] + Rm[2] * bc.qc[2] * R[c * 3 + 2];
          Sv[c] = Rm[0] * bc.qvo[0] * R[c * 3 + 0] + Rm[1] * bc.qvo[1] * R[c * 3 + 1] + Rm[2] * bc.qvo[2] * R[c * 3 + 2];
        }
        double Qrow[9], AtQ[9];
        const double kb_p = -(h * qa_m + dt * qb_m), kb_v = -(h * qb_m + dt * qc_m);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          Qrow[c] = t == 0 ? Sa[c] : (t == 1 ? Sb[c] : 0.0);
          Qrow[3 + c] = t == 0 ? Sb[c] : (t == 1 ? Sc[c] : 0.0);
          Qrow[6 + c] = (t == 2 && c == m) ? qab_m : 0.0;
          AtQ[c] = t == 0 ? Sa[c] : (t == 1 ? dt * Sa[c] + Sb[c] : kb_p * Rcm[c]);
          AtQ[3 + c] = t == 0 ? Sb[c] : (t == 1 ? dt * Sb[c] + Sc[c] : kb_v * Rcm[c]);
          AtQ[6 + c] = Qrow[6 + c];
        }
        double AtQA[9], E345[3];
        bt_row_times_A(AtQ, R, dt, h, AtQA);
        bt_row_times_A(Qrow, R, dt, h, G);
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          G[c] = -G[c];  // row r of E' = column r of E = -(row r of Q A)
          Dn[c] = Qrow[c];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) E345[c] = -AtQ[3 + c];  // E[r][3 + c]
#pragma unroll
        for (int c = 0; c < 9; ++c) Erow[c] = -AtQ[c];      // row r of E (the VO term only enters the p rows: unused there)
        // c_j = (h a_s, dt a_s, 0):  r_j -= A'Q c,  r_{j+1} += Q c
        double rj = 0.0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          rj -= AtQ[c] * (h * as[c]) + AtQ[3 + c] * (dt * as[c]);
          rn += Qrow[c] * (h * as[c]) + Qrow[3 + c] * (dt * as[c]);
        }
        if (vo && t == 0) {  // VO rows p_j - p_{j+1} + Delta, weight Qc = R diag(qvo) R'
          double qd = 0.0;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            AtQA[c] += Sv[c];
            Dn[c] += Sv[c];
            G[c] -= Sv[c];
            qd += Sv[c] * dlt[c];
          }
          rj -= qd;
          rn += qd;
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) D[c] += AtQA[c];
        rhs += rj;
        rhs0 += rj;
        rn0 = rn;
        // v rows of the unmodified blocks -> scratch (multipliers, below)
        if (valid && t == 1) {
          double *gj = fac + (size_t)j * BOX_TFAC + BOX_TFAC_GRAD + m * 28;
#pragma unroll
          for (int c = 0; c < 9; ++c) {
            gj[c] = D[c];
            gj[9 + c] = Erow[c];
            gj[18 + c] = G[c];
          }
          gj[27] = rhs0;
        }
        // fixed components of state j+1: their column of E moves to the right-hand side of state j and is zeroed
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (box_is_active(mask_n, c)) {
            rhs -= E345[c] * box_bound(bc, mask_n, c);
            if (r == 3 + c) {
#pragma unroll
              for (int cc = 0; cc < 9; ++cc) G[cc] = 0.0;
            }
          }
        // fixed components of state j: their row of E feeds the right-hand side of state j+1 and is zeroed
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (box_is_active(mask, c)) {
            rn -= G[3 + c] * box_bound(bc, mask, c);
            G[3 + c] = 0.0;
          }
      }
      if (last && valid && t == 1) {
        double *gj = fac + (size_t)j * BOX_TFAC + BOX_TFAC_GRAD + m * 28;
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          gj[c] = D[c];
          gj[9 + c] = 0.0;
          gj[18 + c] = 0.0;
        }
        gj[27] = rhs0;
      }
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (box_is_active(mask, c)) rhs -= D[3 + c] * box_bound(bc, mask, c);
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (box_is_active(mask, c)) {
          D[3 + c] = 0.0;
          if (r == 3 + c) {
#pragma unroll
            for (int cc = 0; cc < 9; ++cc) D[cc] = 0.0;
            D[3 + c] = 1.0;
            rhs = box_bound(bc, mask, c);
          }
        }
      // S_j = D_j - F_{j-1} F_{j-1}',  rhs_j -= F_{j-1} y_{j-1}   (F_{-1} = 0)
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        double acc = 0.0;
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) acc += Fp[kk] * bt_shfl(Fp[kk], base + c);
        D[c] -= acc;
      }
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) rhs -= Fp[kk] * bt_shfl(yp, base + kk);
      // right-looking Cholesky of S_j fused with y_j = L^-1 rhs_j and F_j = E_j' L_j^-T
      double F[9], ymine = 0.0;
      BtFactor<0>::run(D, G, F, rhs, ymine, status, base, r);
      // factor of this state -> scratch: L (packed lower, row r), F (row r), y
      if (valid) {
        double *fj = fac + (size_t)j * BOX_TFAC;
        const int lo = r * (r + 1) / 2;
#pragma unroll
        for (int c = 0; c < 9; ++c)
          if (c <= r) fj[lo + c] = D[c];
        if (!last) {
#pragma unroll
          for (int c = 0; c < 9; ++c) fj[45 + r * 9 + c] = F[c];
        }
        fj[BOX_TFAC_Y + r] = ymine;
      }
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        Dc[c] = Dn[c];
        Fp[c] = last ? 0.0 : F[c];
      }
      rc = rn;
      rc0 = rn0;
      yp = ymine;
      slot = slot_n;
    }
    __syncwarp();
    // ------------------------------------------------------------------ backward: x_j = L_j^-T (y_j - F_j' x_{j+1})
    double xn = 0.0;
    for (int j = K - 1; j >= 0; --j) {
      double *fj = fac + (size_t)j * BOX_TFAC;
      if (j > 0) bt_prefetch(fj - BOX_TFAC + 16 * r);
      double tt = fj[BOX_TFAC_Y + r];
      if (j + 1 < K) {
#pragma unroll
        for (int kk = 0; kk < 9; ++kk) tt -= fj[45 + kk * 9 + r] * bt_shfl(xn, base + kk);
      }
      double Lcol[9];  // L[kk][r], kk > r; 1 / L[r][r] at kk == r
#pragma unroll
      for (int kk = 0; kk < 9; ++kk) Lcol[kk] = kk >= r ? fj[kk * (kk + 1) / 2 + r] : 1.0;
      double xmine = 0.0;
#pragma unroll
      for (int rr = 8; rr >= 0; --rr) {
        const double xrr = bt_shfl(tt * Lcol[rr], base + rr);  // Lcol[r] of lane r = 1 / L[r][r]
        if (r == rr) xmine = xrr;
        if (r < rr) tt -= Lcol[rr] * xrr;
      }
      if (valid) fj[BOX_TFAC_X + r] = xmine;
      xn = xmine;
      if (j == K - 1) xT_r = xmine;
    }
    __syncwarp();
    // ------------------------------------------------------------------ multipliers and active-set update
    //   gradient of the free cost w.r.t. v_j (rows 3..5 of the block-tridiagonal system, unmodified blocks):
    //   g_j = D0_j x_j + E_j x_{j+1} + E_{j-1}' x_{j-1} - r0_j      -- lane 3 + a evaluates component a
    bool changed = false;
    int na = 0;
    const int q = t == 1 ? m : 0;  // lanes outside the v rows shadow row 3 (their result is not used)
    double xm[9], xj[9], x1[9];
#pragma unroll
    for (int f = 0; f < 9; ++f) {
      xm[f] = 0.0;
      xj[f] = fac[BOX_TFAC_X + f];
    }
    slot = slot0;
    for (int j = 0; j < K; ++j) {
      const double *gj = fac + (size_t)j * BOX_TFAC + BOX_TFAC_GRAD + q * 28;
      const bool more = j + 1 < K;
      if (more && r < 6) bt_prefetch(fac + (size_t)(j + 1) * BOX_TFAC + BOX_TFAC_X + 16 * r);  // x and gradient rows of the next stage
      double g = -gj[27];
#pragma unroll
      for (int f = 0; f < 9; ++f) {
        x1[f] = more ? fac[(size_t)(j + 1) * BOX_TFAC + BOX_TFAC_X + f] : 0.0;
        g += gj[f] * xj[f] + gj[9 + f] * x1[f];
      }
      if (j > 0) {
        const double *gp = gj - BOX_TFAC;
#pragma unroll
        for (int f = 0; f < 9; ++f) g += gp[18 + f] * xm[f];
      }
      uint8_t *mp = bb.act + (size_t)slot * ns + i;
      slot = slot + 1 == dm.NW ? 0 : slot + 1;
      const int mask = *mp;
      const double vq = bt_sel3(q, xj[3], xj[4], xj[5]);
      const double hi_q = bt_sel3(q, bc.hi[0], bc.hi[1], bc.hi[2]), lo_q = bt_sel3(q, bc.lo[0], bc.lo[1], bc.lo[2]);
      bool up = false, dn = false;
      if (mask & (8 << q)) up = -g > 0.0;          // multiplier of the upper bound stays positive
      else if (mask & (1 << q)) dn = g > 0.0;
      else if (vq > hi_q) up = true;
      else if (vq < lo_q) dn = true;
      const unsigned bu = __ballot_sync(0xffffffffu, up), bd = __ballot_sync(0xffffffffu, dn);
      const int nm = (int)((bd >> (base + 3)) & 7u) | ((int)((bu >> (base + 3)) & 7u) << 3);
      na += __popc((unsigned)nm);
      __syncwarp();  // every lane of the team has read the mask before lane 0 replaces it
      if (nm != mask) {
        if (!changed) first_changed = j;
        changed = true;
        if (valid && r == 0) *mp = (uint8_t)nm;
      }
#pragma unroll
      for (int f = 0; f < 9; ++f) {
        xm[f] = xj[f];
        xj[f] = x1[f];
      }
    }
    __syncwarp();
    if (!done) nact = na;
    if (!changed) done = true;
  }
  if (!done) status |= ST_QP_MAXITER;
  if (valid && r == 0) {
    bb.iters[i] = iters;
    bb.nactive[i] = nact;
  }
  return status;
}

#endif  // __CUDACC__

// Prior of the constrained solve: marginalizeQP(T-N) as one stage of the unconstrained covariance-form sweep (the first
// half of mhe_solve_box), then M = Pa^-1 (symmetrised) and m = M xa to tb.prior.  One thread per instance.
template <typename T, typename Math = DefaultMath<T>>
DEKF_HD int box_team_prior(const MheConst<T> &c, const Dims &dm, const Buffers<T> &b, const BoxBuffers &bb, const BoxTeamBuffers &tb,
                           int Tk, int i) {
  double Pa[81], xa[9];
  int k0;
  box_prepare<T, Math>(c, dm, b, Tk, i, Pa, xa, k0);
  double M[81], mv[9];
  int status = box_prior_info(Pa, xa, M, mv);
  double *pri = tb.prior + (size_t)i * BOX_PRIOR;
  for (int rr = 0; rr < 9; ++rr)
    for (int cc = 0; cc < 9; ++cc) pri[rr * 9 + cc] = 0.5 * (M[rr * 9 + cc] + M[cc * 9 + rr]);
  for (int f = 0; f < 9; ++f) pri[81 + f] = mv[f];
  bb.act[(size_t)(Tk % dm.NW) * dm.ns + i] = 0;  // the new state starts free (warm start keeps the older masks)
  return status;
}

}  // namespace dekf
