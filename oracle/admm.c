/* CPU oracle: OSQP-style ADMM.  TEST INFRASTRUCTURE ONLY.
 *
 * OSQP / osqp-eigen are NOT under /root/reference (find_package(OsqpEigen REQUIRED), unversioned,
 * src/decentral_legged_est/CMakeLists.txt:13; README.md:18 points at osqp.org; API names used at
 * DecentralEst.cpp:204-217 indicate osqp 0.6.x + osqp-eigen 0.7/0.8 -- inference, no pin).  This
 * file restates the PUBLISHED algorithm (Stellato et al., "OSQP: an operator splitting solver for
 * quadratic programs", Math. Prog. Comp. 2020, Alg. 1 + sec. 5: Ruiz equilibration, per-constraint
 * rho, adaptive rho, unscaled termination test) with the settings the reference's call sites pass
 * (MheSrb.cpp:278-293 cold setup every step, :345-346 solve + getSolution unconditionally):
 *   scaling=10, check_termination=25, adaptive_rho_tolerance=5, linsys = sparse LDL' of the
 *   quasi-definite KKT [[P+sigma I, A'],[A, -diag(1/rho)]].
 * OSQP picks the adaptive-rho interval from measured setup time (adaptive_rho_interval=0); here
 * it is the deterministic 25 (ORC_ADMM_RHO_INTERVAL overrides).  "parity unpinned". */
#include "oracle.h"
#include "la.h"
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-04
#define MAX_SCALING 1e+04
#define RHO_MIN 1e-06
#define RHO_MAX 1e06
#define RHO_EQ_OVER_RHO_INEQ 1e03
#define RHO_TOL 1e-04
#define SCALING_ITERS 10
#define CHECK_TERMINATION 25
#define ADAPTIVE_RHO_TOLERANCE 5.0

typedef struct {
  int n, m;      /* rows n, cols m */
  int *p, *i;    /* column pointers (m+1), row indices */
  double *x;
} csc_t;

static csc_t csc_from_dense(const double *D, int rows, int cols, int upper_only) {
  csc_t c;
  c.n = rows;
  c.m = cols;
  c.p = (int *)calloc((size_t)cols + 1, sizeof(int));
  int nnz = 0;
  for (int j = 0; j < cols; ++j)
    for (int i = 0; i < rows; ++i)
      if (D[(size_t)i * cols + j] != 0.0 && (!upper_only || i <= j)) nnz++;
  c.i = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
  c.x = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  nnz = 0;
  for (int j = 0; j < cols; ++j) {
    c.p[j] = nnz;
    for (int i = 0; i < rows; ++i)
      if (D[(size_t)i * cols + j] != 0.0 && (!upper_only || i <= j)) {
        c.i[nnz] = i;
        c.x[nnz] = D[(size_t)i * cols + j];
        nnz++;
      }
  }
  c.p[cols] = nnz;
  return c;
}
static void csc_free(csc_t *c) {
  free(c->p);
  free(c->i);
  free(c->x);
}

/* y = P x with P symmetric stored upper */
static void sym_mv(const csc_t *P, const double *x, double *y) {
  for (int i = 0; i < P->n; ++i) y[i] = 0.0;
  for (int j = 0; j < P->m; ++j)
    for (int k = P->p[j]; k < P->p[j + 1]; ++k) {
      int i = P->i[k];
      y[i] += P->x[k] * x[j];
      if (i != j) y[j] += P->x[k] * x[i];
    }
}
static void mat_mv(const csc_t *A, const double *x, double *y) { /* y = A x */
  for (int i = 0; i < A->n; ++i) y[i] = 0.0;
  for (int j = 0; j < A->m; ++j)
    for (int k = A->p[j]; k < A->p[j + 1]; ++k) y[A->i[k]] += A->x[k] * x[j];
}
static void mat_tmv(const csc_t *A, const double *x, double *y) { /* y = A' x */
  for (int j = 0; j < A->m; ++j) {
    double s = 0.0;
    for (int k = A->p[j]; k < A->p[j + 1]; ++k) s += A->x[k] * x[A->i[k]];
    y[j] = s;
  }
}
static double norm_inf(const double *x, int n) {
  double m = 0.0;
  for (int i = 0; i < n; ++i)
    if (fabs(x[i]) > m) m = fabs(x[i]);
  return m;
}
static double scaled_norm_inf(const double *s, const double *x, int n) {
  double m = 0.0;
  for (int i = 0; i < n; ++i)
    if (fabs(s[i] * x[i]) > m) m = fabs(s[i] * x[i]);
  return m;
}

/* ---------------------------------------------------------------- ordering (minimum degree) */
static int popcount_and(const uint64_t *a, const uint64_t *mask, int nw) {
  int c = 0;
  for (int w = 0; w < nw; ++w) c += __builtin_popcountll(a[w] & mask[w]);
  return c;
}

/* pattern: upper-triangular KKT in CSC (dimension N).  perm[k] = original index of k-th pivot. */
static void min_degree_order(int N, const int *Kp, const int *Ki, int *perm) {
  int nw = (N + 63) / 64;
  uint64_t *adj = (uint64_t *)calloc((size_t)N * nw, sizeof(uint64_t));
  uint64_t *active = (uint64_t *)calloc((size_t)nw, sizeof(uint64_t));
  int *deg = (int *)malloc(sizeof(int) * (size_t)N);
  for (int j = 0; j < N; ++j) {
    active[j / 64] |= (uint64_t)1 << (j % 64);
    for (int k = Kp[j]; k < Kp[j + 1]; ++k) {
      int i = Ki[k];
      if (i == j) continue;
      adj[(size_t)i * nw + j / 64] |= (uint64_t)1 << (j % 64);
      adj[(size_t)j * nw + i / 64] |= (uint64_t)1 << (i % 64);
    }
  }
  for (int j = 0; j < N; ++j) deg[j] = popcount_and(&adj[(size_t)j * nw], active, nw);
  for (int step = 0; step < N; ++step) {
    int best = -1, bd = 1 << 30;
    for (int j = 0; j < N; ++j)
      if ((active[j / 64] >> (j % 64)) & 1)
        if (deg[j] < bd) {
          bd = deg[j];
          best = j;
        }
    perm[step] = best;
    active[best / 64] &= ~((uint64_t)1 << (best % 64));
    const uint64_t *ab = &adj[(size_t)best * nw];
    for (int w = 0; w < nw; ++w) {
      uint64_t bits = ab[w] & active[w];
      while (bits) {
        int b = __builtin_ctzll(bits);
        bits &= bits - 1;
        int i = w * 64 + b;
        uint64_t *ai = &adj[(size_t)i * nw];
        for (int w2 = 0; w2 < nw; ++w2) ai[w2] |= ab[w2];
        ai[i / 64] &= ~((uint64_t)1 << (i % 64));
        deg[i] = popcount_and(ai, active, nw);
      }
    }
  }
  free(adj);
  free(active);
  free(deg);
}

/* ---------------------------------------------------------------- sparse LDL' (up-looking) */
typedef struct {
  int N;
  int *perm, *iperm;
  /* permuted upper-triangular KKT pattern + map from (P,A,diag) entries to its values */
  int *Kp, *Ki;
  double *Kx;
  int *etree, *Lnz, *Lp, *Li;
  double *Lx, *D, *Dinv;
  /* work */
  int *flag, *pattern, *lnz_cur;
  double *y;
  double *bp;
} ldl_t;

static void ldl_symbolic(ldl_t *f) {
  int N = f->N;
  f->etree = (int *)malloc(sizeof(int) * (size_t)N);
  f->Lnz = (int *)calloc((size_t)N, sizeof(int));
  f->flag = (int *)malloc(sizeof(int) * (size_t)N);
  for (int k = 0; k < N; ++k) {
    f->etree[k] = -1;
    f->flag[k] = k;
    for (int p = f->Kp[k]; p < f->Kp[k + 1]; ++p) {
      int i = f->Ki[p];
      if (i >= k) continue;
      for (; f->flag[i] != k; i = f->etree[i]) {
        if (f->etree[i] == -1) f->etree[i] = k;
        f->Lnz[i]++;
        f->flag[i] = k;
      }
    }
  }
  f->Lp = (int *)malloc(sizeof(int) * ((size_t)N + 1));
  f->Lp[0] = 0;
  for (int k = 0; k < N; ++k) f->Lp[k + 1] = f->Lp[k] + f->Lnz[k];
  int lnz = f->Lp[N];
  f->Li = (int *)malloc(sizeof(int) * (size_t)(lnz > 0 ? lnz : 1));
  f->Lx = (double *)malloc(sizeof(double) * (size_t)(lnz > 0 ? lnz : 1));
  f->D = (double *)malloc(sizeof(double) * (size_t)N);
  f->Dinv = (double *)malloc(sizeof(double) * (size_t)N);
  f->pattern = (int *)malloc(sizeof(int) * (size_t)N);
  f->lnz_cur = (int *)malloc(sizeof(int) * (size_t)N);
  f->y = (double *)calloc((size_t)N, sizeof(double));
  f->bp = (double *)malloc(sizeof(double) * (size_t)N);
}

static int ldl_numeric(ldl_t *f) {
  int N = f->N;
  for (int k = 0; k < N; ++k) f->lnz_cur[k] = 0;
  for (int k = 0; k < N; ++k) {
    int top = N;
    f->flag[k] = k;
    f->y[k] = 0.0;
    for (int p = f->Kp[k]; p < f->Kp[k + 1]; ++p) {
      int i = f->Ki[p];
      if (i > k) continue;
      f->y[i] += f->Kx[p];
      int len = 0;
      for (; f->flag[i] != k; i = f->etree[i]) {
        f->pattern[len++] = i;
        f->flag[i] = k;
      }
      while (len > 0) f->pattern[--top] = f->pattern[--len];
    }
    double dk = f->y[k];
    f->y[k] = 0.0;
    for (; top < N; ++top) {
      int i = f->pattern[top];
      double yi = f->y[i];
      f->y[i] = 0.0;
      int p2 = f->Lp[i] + f->lnz_cur[i];
      for (int p = f->Lp[i]; p < p2; ++p) f->y[f->Li[p]] -= f->Lx[p] * yi;
      double lki = yi * f->Dinv[i];
      dk -= lki * yi;
      f->Li[p2] = k;
      f->Lx[p2] = lki;
      f->lnz_cur[i]++;
    }
    if (dk == 0.0) return -1;
    f->D[k] = dk;
    f->Dinv[k] = 1.0 / dk;
  }
  return 0;
}

static void ldl_solve(const ldl_t *f, double *b) {
  int N = f->N;
  double *x = f->bp;
  for (int k = 0; k < N; ++k) x[k] = b[f->perm[k]];
  for (int j = 0; j < N; ++j) {
    double xj = x[j];
    for (int p = f->Lp[j]; p < f->Lp[j + 1]; ++p) x[f->Li[p]] -= f->Lx[p] * xj;
  }
  for (int j = 0; j < N; ++j) x[j] *= f->Dinv[j];
  for (int j = N - 1; j >= 0; --j) {
    double xj = x[j];
    for (int p = f->Lp[j]; p < f->Lp[j + 1]; ++p) xj -= f->Lx[p] * x[f->Li[p]];
    x[j] = xj;
  }
  for (int k = 0; k < N; ++k) b[f->perm[k]] = x[k];
}

static void ldl_free(ldl_t *f) {
  free(f->iperm);
  free(f->Kp);
  free(f->Ki);
  free(f->Kx);
  free(f->etree);
  free(f->Lnz);
  free(f->Lp);
  free(f->Li);
  free(f->Lx);
  free(f->D);
  free(f->Dinv);
  free(f->flag);
  free(f->pattern);
  free(f->lnz_cur);
  free(f->y);
  free(f->bp);
}

/* ordering cache keyed by (n, m, nnzP, nnzA): the pattern is identical from step to step. */
static struct {
  int n, m, nnzP, nnzA;
  int *perm;
} g_cache = {0, 0, 0, 0, NULL};
#if defined(__GNUC__)
#define ORC_TLS __thread
#else
#define ORC_TLS
#endif
static ORC_TLS int *tls_perm = NULL;
static ORC_TLS int tls_key[4] = {0, 0, 0, 0};

/* Assemble permuted upper-triangular KKT = [[P+sigma I, A'],[A, -diag(1/rho)]] */
static void kkt_build(ldl_t *f, const csc_t *P, const csc_t *A, double sigma, const double *rho_vec,
                      int first) {
  int n = P->n, m = A->n, N = n + m;
  if (first) {
    f->N = N;
    /* unpermuted pattern (upper): col j<n: P(:,j) upper + diag; col n+i: A(i,:) entries as rows j, diag */
    int nnz = P->p[n] + n + A->p[A->m] + m;
    int *Tp = (int *)calloc((size_t)N + 1, sizeof(int));
    int *Ti = (int *)malloc(sizeof(int) * (size_t)nnz);
    /* count */
    int *cnt = (int *)calloc((size_t)N, sizeof(int));
    for (int j = 0; j < n; ++j) {
      int has_diag = 0;
      for (int k = P->p[j]; k < P->p[j + 1]; ++k) {
        cnt[j]++;
        if (P->i[k] == j) has_diag = 1;
      }
      if (!has_diag) cnt[j]++;
    }
    for (int j = 0; j < A->m; ++j)
      for (int k = A->p[j]; k < A->p[j + 1]; ++k) cnt[n + A->i[k]]++;
    for (int i = 0; i < m; ++i) cnt[n + i]++;
    for (int j = 0; j < N; ++j) Tp[j + 1] = Tp[j] + cnt[j];
    memset(cnt, 0, sizeof(int) * (size_t)N);
    for (int j = 0; j < n; ++j) {
      int has_diag = 0;
      for (int k = P->p[j]; k < P->p[j + 1]; ++k) {
        Ti[Tp[j] + cnt[j]++] = P->i[k];
        if (P->i[k] == j) has_diag = 1;
      }
      if (!has_diag) Ti[Tp[j] + cnt[j]++] = j;
    }
    for (int j = 0; j < A->m; ++j)
      for (int k = A->p[j]; k < A->p[j + 1]; ++k) {
        int c = n + A->i[k];
        Ti[Tp[c] + cnt[c]++] = j;
      }
    for (int i = 0; i < m; ++i) Ti[Tp[n + i] + cnt[n + i]++] = n + i;
    int key[4] = {n, m, P->p[n], A->p[A->m]};
    if (!(tls_perm && memcmp(key, tls_key, sizeof(key)) == 0)) {
      free(tls_perm);
      tls_perm = (int *)malloc(sizeof(int) * (size_t)N);
      min_degree_order(N, Tp, Ti, tls_perm);
      memcpy(tls_key, key, sizeof(key));
    }
    f->perm = tls_perm;
    f->iperm = (int *)malloc(sizeof(int) * (size_t)N);
    for (int k = 0; k < N; ++k) f->iperm[f->perm[k]] = k;
    /* permuted pattern: entry (i,j) -> (min(pi,pj), max(pi,pj)) */
    memset(cnt, 0, sizeof(int) * (size_t)N);
    for (int j = 0; j < N; ++j)
      for (int k = Tp[j]; k < Tp[j + 1]; ++k) {
        int a = f->iperm[Ti[k]], b = f->iperm[j];
        cnt[a > b ? a : b]++;
      }
    f->Kp = (int *)calloc((size_t)N + 1, sizeof(int));
    for (int j = 0; j < N; ++j) f->Kp[j + 1] = f->Kp[j] + cnt[j];
    f->Ki = (int *)malloc(sizeof(int) * (size_t)nnz);
    f->Kx = (double *)malloc(sizeof(double) * (size_t)nnz);
    free(Tp);
    free(Ti);
    free(cnt);
  }
  /* fill values (pattern order: recomputed the same way each time) */
  int N2 = f->N;
  int *cur = (int *)calloc((size_t)N2, sizeof(int));
#define PUT(ii, jj, val)                                     \
  do {                                                       \
    int a_ = f->iperm[ii], b_ = f->iperm[jj];                \
    int c_ = a_ > b_ ? a_ : b_, r_ = a_ > b_ ? b_ : a_;      \
    f->Ki[f->Kp[c_] + cur[c_]] = r_;                         \
    f->Kx[f->Kp[c_] + cur[c_]] = (val);                      \
    cur[c_]++;                                               \
  } while (0)
  for (int j = 0; j < n; ++j) {
    int has_diag = 0;
    for (int k = P->p[j]; k < P->p[j + 1]; ++k) {
      if (P->i[k] == j) {
        has_diag = 1;
        PUT(j, j, P->x[k] + sigma);
      } else {
        PUT(P->i[k], j, P->x[k]);
      }
    }
    if (!has_diag) PUT(j, j, sigma);
  }
  for (int j = 0; j < A->m; ++j)
    for (int k = A->p[j]; k < A->p[j + 1]; ++k) PUT(j, n + A->i[k], A->x[k]);
  for (int i = 0; i < m; ++i) PUT(n + i, n + i, -1.0 / rho_vec[i]);
#undef PUT
  free(cur);
  (void)g_cache;
}

/* ---------------------------------------------------------------- OSQP-style solve */
static double limit_scaling(double v) {
  v = v < MIN_SCALING ? 1.0 : v;
  v = v > MAX_SCALING ? MAX_SCALING : v;
  return v;
}

static void set_rho_vec(int m, const double *l, const double *u, double rho, double *rho_vec, int *ctype) {
  for (int i = 0; i < m; ++i) {
    if (l[i] < -OSQP_INFTY * MIN_SCALING && u[i] > OSQP_INFTY * MIN_SCALING) {
      ctype[i] = -1;
      rho_vec[i] = RHO_MIN;
    } else if (u[i] - l[i] < RHO_TOL) {
      ctype[i] = 1;
      rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * rho;
    } else {
      ctype[i] = 0;
      rho_vec[i] = rho;
    }
  }
}

static csc_t csc_from_triplets(int rows, int cols, int nnz, const int *ti, const int *tj, const double *tx) {
  csc_t c;
  c.n = rows;
  c.m = cols;
  c.p = (int *)calloc((size_t)cols + 1, sizeof(int));
  c.i = (int *)malloc(sizeof(int) * (size_t)(nnz > 0 ? nnz : 1));
  c.x = (double *)malloc(sizeof(double) * (size_t)(nnz > 0 ? nnz : 1));
  for (int k = 0; k < nnz; ++k) c.p[tj[k] + 1]++;
  for (int j = 0; j < cols; ++j) c.p[j + 1] += c.p[j];
  int *cur = (int *)malloc(sizeof(int) * (size_t)(cols > 0 ? cols : 1));
  for (int j = 0; j < cols; ++j) cur[j] = c.p[j];
  for (int k = 0; k < nnz; ++k) {
    int d = cur[tj[k]]++;
    c.i[d] = ti[k];
    c.x[d] = tx[k];
  }
  free(cur);
  return c;
}

static int admm_core(csc_t P, csc_t A, const double *gd, const double *ld, const double *ud,
                     const orc_params *prm, double *z_out, int *iters_out);

/* dense entry point (H symmetric n x n, A m x n, both row-major) */
int orc_admm_solve(int n, int m, const double *Hd, const double *gd, const double *Ad, const double *ld,
                   const double *ud, const orc_params *prm, double *z_out, int *iters_out) {
  csc_t P = csc_from_dense(Hd, n, n, 1);
  csc_t A = csc_from_dense(Ad, m, n, 0);
  return admm_core(P, A, gd, ld, ud, prm, z_out, iters_out);
}

/* triplet entry point: P entries must be the UPPER triangle (i <= j), no duplicates */
int orc_admm_solve_triplets(int n, int m, int nnzP, const int *Pi, const int *Pj, const double *Px,
                            int nnzA, const int *Ai, const int *Aj, const double *Ax, const double *gd,
                            const double *ld, const double *ud, const orc_params *prm, double *z_out,
                            int *iters_out) {
  csc_t P = csc_from_triplets(n, n, nnzP, Pi, Pj, Px);
  csc_t A = csc_from_triplets(m, n, nnzA, Ai, Aj, Ax);
  return admm_core(P, A, gd, ld, ud, prm, z_out, iters_out);
}

static int admm_core(csc_t P, csc_t A, const double *gd, const double *ld, const double *ud,
                     const orc_params *prm, double *z_out, int *iters_out) {
  int n = P.n, m = A.n;
  double *q = (double *)malloc(sizeof(double) * (size_t)n);
  double *l = (double *)malloc(sizeof(double) * (size_t)m), *u = (double *)malloc(sizeof(double) * (size_t)m);
  memcpy(q, gd, sizeof(double) * (size_t)n);
  for (int i = 0; i < m; ++i) { /* OSQP clips bounds at +-OSQP_INFTY */
    l[i] = ld[i] < -OSQP_INFTY ? -OSQP_INFTY : ld[i];
    u[i] = ud[i] > OSQP_INFTY ? OSQP_INFTY : ud[i];
  }
  /* ---- Ruiz equilibration (scaling = 10) */
  double *D = (double *)malloc(sizeof(double) * (size_t)n), *E = (double *)malloc(sizeof(double) * (size_t)m);
  double *Dt = (double *)malloc(sizeof(double) * (size_t)n), *Et = (double *)malloc(sizeof(double) * (size_t)m);
  double c = 1.0;
  for (int i = 0; i < n; ++i) D[i] = 1.0;
  for (int i = 0; i < m; ++i) E[i] = 1.0;
  for (int it = 0; it < SCALING_ITERS; ++it) {
    for (int i = 0; i < n; ++i) Dt[i] = 0.0;
    for (int i = 0; i < m; ++i) Et[i] = 0.0;
    for (int j = 0; j < n; ++j)
      for (int k = P.p[j]; k < P.p[j + 1]; ++k) {
        double a = fabs(P.x[k]);
        int i = P.i[k];
        if (a > Dt[j]) Dt[j] = a;
        if (a > Dt[i]) Dt[i] = a;
      }
    for (int j = 0; j < n; ++j)
      for (int k = A.p[j]; k < A.p[j + 1]; ++k) {
        double a = fabs(A.x[k]);
        if (a > Dt[j]) Dt[j] = a;
        if (a > Et[A.i[k]]) Et[A.i[k]] = a;
      }
    for (int i = 0; i < n; ++i) Dt[i] = 1.0 / sqrt(limit_scaling(Dt[i]));
    for (int i = 0; i < m; ++i) Et[i] = 1.0 / sqrt(limit_scaling(Et[i]));
    for (int j = 0; j < n; ++j)
      for (int k = P.p[j]; k < P.p[j + 1]; ++k) P.x[k] *= Dt[P.i[k]] * Dt[j];
    for (int j = 0; j < n; ++j)
      for (int k = A.p[j]; k < A.p[j + 1]; ++k) A.x[k] *= Et[A.i[k]] * Dt[j];
    for (int i = 0; i < n; ++i) {
      q[i] *= Dt[i];
      D[i] *= Dt[i];
    }
    for (int i = 0; i < m; ++i) E[i] *= Et[i];
    /* cost normalisation */
    for (int i = 0; i < n; ++i) Dt[i] = 0.0;
    for (int j = 0; j < n; ++j)
      for (int k = P.p[j]; k < P.p[j + 1]; ++k) {
        double a = fabs(P.x[k]);
        int i = P.i[k];
        if (a > Dt[j]) Dt[j] = a;
        if (a > Dt[i]) Dt[i] = a;
      }
    double c_temp = 0.0;
    for (int i = 0; i < n; ++i) c_temp += Dt[i];
    c_temp /= n;
    double inf_norm_q = norm_inf(q, n);
    inf_norm_q = inf_norm_q < MIN_SCALING ? 1.0 : (inf_norm_q > MAX_SCALING ? MAX_SCALING : inf_norm_q);
    c_temp = c_temp > inf_norm_q ? c_temp : inf_norm_q;
    c_temp = limit_scaling(c_temp);
    c_temp = 1.0 / c_temp;
    for (int k = 0; k < P.p[n]; ++k) P.x[k] *= c_temp;
    for (int i = 0; i < n; ++i) q[i] *= c_temp;
    c *= c_temp;
  }
  double *Dinv = Dt, *Einv = Et;
  for (int i = 0; i < n; ++i) Dinv[i] = 1.0 / D[i];
  for (int i = 0; i < m; ++i) {
    Einv[i] = 1.0 / E[i];
    l[i] *= E[i];
    u[i] *= E[i];
  }
  double cinv = 1.0 / c;

  /* ---- setup linear system */
  double rho = prm->rho, sigma = prm->sigma, alpha = prm->alpha;
  double eps_abs = prm->abs_tol, eps_rel = prm->relative_tol;
  int max_iter = prm->max_qp_iter;
  int rho_interval = CHECK_TERMINATION;
  {
    const char *e = getenv("ORC_ADMM_RHO_INTERVAL");
    if (e) rho_interval = atoi(e);
  }
  double *rho_vec = (double *)malloc(sizeof(double) * (size_t)m);
  int *ctype = (int *)malloc(sizeof(int) * (size_t)m);
  set_rho_vec(m, l, u, rho, rho_vec, ctype);
  ldl_t F;
  memset(&F, 0, sizeof(F));
  kkt_build(&F, &P, &A, sigma, rho_vec, 1);
  ldl_symbolic(&F);
  ldl_numeric(&F);

  int N = n + m;
  double *x = (double *)calloc((size_t)n, sizeof(double)), *z = (double *)calloc((size_t)m, sizeof(double));
  double *y = (double *)calloc((size_t)m, sizeof(double));
  double *x_prev = (double *)calloc((size_t)n, sizeof(double)), *z_prev = (double *)calloc((size_t)m, sizeof(double));
  double *xz = (double *)calloc((size_t)N, sizeof(double));
  double *Ax = (double *)calloc((size_t)m, sizeof(double)), *Px = (double *)calloc((size_t)n, sizeof(double));
  double *Aty = (double *)calloc((size_t)n, sizeof(double)), *tmpn = (double *)calloc((size_t)n, sizeof(double));
  double *tmpm = (double *)calloc((size_t)m, sizeof(double));
  int iter, solved = 0;
  for (iter = 1; iter <= max_iter; ++iter) {
    memcpy(x_prev, x, sizeof(double) * (size_t)n);
    memcpy(z_prev, z, sizeof(double) * (size_t)m);
    for (int i = 0; i < n; ++i) xz[i] = sigma * x_prev[i] - q[i];
    for (int i = 0; i < m; ++i) xz[n + i] = z_prev[i] - y[i] / rho_vec[i];
    ldl_solve(&F, xz);
    for (int i = 0; i < m; ++i) xz[n + i] = z_prev[i] + (xz[n + i] - y[i]) / rho_vec[i]; /* z_tilde */
    for (int i = 0; i < n; ++i) x[i] = alpha * xz[i] + (1.0 - alpha) * x_prev[i];
    for (int i = 0; i < m; ++i) {
      double zt = alpha * xz[n + i] + (1.0 - alpha) * z_prev[i];
      double v = zt + y[i] / rho_vec[i];
      v = v < l[i] ? l[i] : (v > u[i] ? u[i] : v);
      z[i] = v;
      y[i] += rho_vec[i] * (zt - v);
    }
    int check = (iter % CHECK_TERMINATION == 0);
    int adapt = prm->adapt_rho && rho_interval > 0 && (iter % rho_interval == 0);
    if (check || adapt) {
      mat_mv(&A, x, Ax);
      sym_mv(&P, x, Px);
      mat_tmv(&A, y, Aty);
      for (int i = 0; i < m; ++i) tmpm[i] = Ax[i] - z[i];
      for (int i = 0; i < n; ++i) tmpn[i] = Px[i] + q[i] + Aty[i];
      if (check) {
        double pri_res = scaled_norm_inf(Einv, tmpm, m);
        double dua_res = cinv * scaled_norm_inf(Dinv, tmpn, n);
        double np1 = scaled_norm_inf(Einv, z, m), np2 = scaled_norm_inf(Einv, Ax, m);
        double eps_prim = eps_abs + eps_rel * (np1 > np2 ? np1 : np2);
        double nd1 = scaled_norm_inf(Dinv, q, n), nd2 = scaled_norm_inf(Dinv, Aty, n), nd3 = scaled_norm_inf(Dinv, Px, n);
        double ndm = nd1 > nd2 ? nd1 : nd2;
        ndm = ndm > nd3 ? ndm : nd3;
        double eps_dual = eps_abs + eps_rel * cinv * ndm;
        if (pri_res < eps_prim && dua_res < eps_dual) {
          solved = 1;
          break;
        }
      }
      if (adapt) {
        double pri = norm_inf(tmpm, m), dua = norm_inf(tmpn, n);
        double pn = norm_inf(z, m), pn2 = norm_inf(Ax, m);
        pn = pn > pn2 ? pn : pn2;
        double dn = norm_inf(q, n), dn2 = norm_inf(Aty, n), dn3 = norm_inf(Px, n);
        dn = dn > dn2 ? dn : dn2;
        dn = dn > dn3 ? dn : dn3;
        pri /= (pn + 1e-10);
        dua /= (dn + 1e-10);
        double rho_new = rho * sqrt(pri / (dua + 1e-10));
        rho_new = rho_new < RHO_MIN ? RHO_MIN : (rho_new > RHO_MAX ? RHO_MAX : rho_new);
        if (rho_new > rho * ADAPTIVE_RHO_TOLERANCE || rho_new < rho / ADAPTIVE_RHO_TOLERANCE) {
          rho = rho_new;
          for (int i = 0; i < m; ++i)
            rho_vec[i] = ctype[i] == -1 ? RHO_MIN : (ctype[i] == 1 ? RHO_EQ_OVER_RHO_INEQ * rho : rho);
          kkt_build(&F, &P, &A, sigma, rho_vec, 0);
          ldl_numeric(&F);
        }
      }
    }
  }
  if (iter > max_iter) iter = max_iter;
  for (int i = 0; i < n; ++i) z_out[i] = D[i] * x[i]; /* unscale_solution */
  if (iters_out) *iters_out = iter;

  ldl_free(&F);
  csc_free(&P);
  csc_free(&A);
  free(q);
  free(l);
  free(u);
  free(D);
  free(E);
  free(Dt);
  free(Et);
  free(rho_vec);
  free(ctype);
  free(x);
  free(z);
  free(y);
  free(x_prev);
  free(z_prev);
  free(xz);
  free(Ax);
  free(Px);
  free(Aty);
  free(tmpn);
  free(tmpm);
  return solved ? 0 : 1;
}
