/* Dense row-major helpers for the CPU oracle.  TEST INFRASTRUCTURE ONLY. */
#include "la.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

void la_mm(double *C, const double *A, const double *B, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[i * k + p] * B[p * n + j];
      C[i * n + j] = s;
    }
}

void la_mmt(double *C, const double *A, const double *B, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[i * k + p] * B[j * k + p];
      C[i * n + j] = s;
    }
}

void la_mtm(double *C, const double *A, const double *B, int m, int k, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) {
      double s = 0.0;
      for (int p = 0; p < k; ++p) s += A[p * m + i] * B[p * n + j];
      C[i * n + j] = s;
    }
}

void la_transpose(double *At, const double *A, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) At[j * m + i] = A[i * n + j];
}

void la_zero(double *A, int n) { memset(A, 0, sizeof(double) * (size_t)n); }

void la_eye(double *A, int n) {
  la_zero(A, n * n);
  for (int i = 0; i < n; ++i) A[i * n + i] = 1.0;
}

void la_copy(double *dst, const double *src, int n) { memcpy(dst, src, sizeof(double) * (size_t)n); }

void la_set_block(double *dst, int ld, int r0, int c0, const double *src, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) dst[(r0 + i) * ld + c0 + j] = src[i * n + j];
}

void la_get_block(double *dst, const double *src, int ld, int r0, int c0, int m, int n) {
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < n; ++j) dst[i * n + j] = src[(r0 + i) * ld + c0 + j];
}

int la_inverse(double *Ainv, const double *A, int n) {
  double *LU = (double *)malloc(sizeof(double) * (size_t)n * n);
  int *piv = (int *)malloc(sizeof(int) * (size_t)n);
  la_copy(LU, A, n * n);
  for (int i = 0; i < n; ++i) piv[i] = i;
  int rc = 0;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(LU[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(LU[i * n + k]) > best) {
        best = fabs(LU[i * n + k]);
        p = i;
      }
    if (best == 0.0) {
      rc = -1;
      break;
    }
    if (p != k) {
      for (int j = 0; j < n; ++j) {
        double t = LU[k * n + j];
        LU[k * n + j] = LU[p * n + j];
        LU[p * n + j] = t;
      }
      int t = piv[k];
      piv[k] = piv[p];
      piv[p] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      LU[i * n + k] /= LU[k * n + k];
      double l = LU[i * n + k];
      for (int j = k + 1; j < n; ++j) LU[i * n + j] -= l * LU[k * n + j];
    }
  }
  if (rc == 0) {
    /* solve for each unit vector */
    double *col = (double *)malloc(sizeof(double) * (size_t)n);
    for (int c = 0; c < n; ++c) {
      for (int i = 0; i < n; ++i) col[i] = (piv[i] == c) ? 1.0 : 0.0;
      for (int i = 0; i < n; ++i) {
        double s = col[i];
        for (int j = 0; j < i; ++j) s -= LU[i * n + j] * col[j];
        col[i] = s;
      }
      for (int i = n - 1; i >= 0; --i) {
        double s = col[i];
        for (int j = i + 1; j < n; ++j) s -= LU[i * n + j] * col[j];
        col[i] = s / LU[i * n + i];
      }
      for (int i = 0; i < n; ++i) Ainv[i * n + c] = col[i];
    }
    free(col);
  }
  free(LU);
  free(piv);
  return rc;
}

static int chol_factor(double *A, int n, int bw) {
  /* lower Cholesky in place; bw<0 means dense */
  for (int j = 0; j < n; ++j) {
    int k0 = (bw < 0) ? 0 : (j - bw > 0 ? j - bw : 0);
    double d = A[j * n + j];
    for (int k = k0; k < j; ++k) d -= A[j * n + k] * A[j * n + k];
    if (!(d > 0.0)) return -1;
    d = sqrt(d);
    A[j * n + j] = d;
    int i1 = (bw < 0) ? n : (j + bw + 1 < n ? j + bw + 1 : n);
    for (int i = j + 1; i < i1; ++i) {
      int kk0 = (bw < 0) ? 0 : (i - bw > 0 ? i - bw : 0);
      if (kk0 < k0) kk0 = k0;
      double s = A[i * n + j];
      for (int k = kk0; k < j; ++k) s -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = s / d;
    }
  }
  return 0;
}

static void chol_subst(const double *L, double *b, int n, int bw) {
  for (int i = 0; i < n; ++i) {
    int k0 = (bw < 0) ? 0 : (i - bw > 0 ? i - bw : 0);
    double s = b[i];
    for (int k = k0; k < i; ++k) s -= L[i * n + k] * b[k];
    b[i] = s / L[i * n + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    int k1 = (bw < 0) ? n : (i + bw + 1 < n ? i + bw + 1 : n);
    double s = b[i];
    for (int k = i + 1; k < k1; ++k) s -= L[k * n + i] * b[k];
    b[i] = s / L[i * n + i];
  }
}

int la_spd_inverse(double *Ainv, const double *A, int n) {
  double *L = (double *)malloc(sizeof(double) * (size_t)n * n);
  double *col = (double *)malloc(sizeof(double) * (size_t)n);
  la_copy(L, A, n * n);
  int rc = chol_factor(L, n, -1);
  if (rc == 0) {
    for (int c = 0; c < n; ++c) {
      for (int i = 0; i < n; ++i) col[i] = (i == c) ? 1.0 : 0.0;
      chol_subst(L, col, n, -1);
      for (int i = 0; i < n; ++i) Ainv[i * n + c] = col[i];
    }
  }
  free(L);
  free(col);
  return rc;
}

int la_chol_solve(double *A, double *b, int n) {
  int rc = chol_factor(A, n, -1);
  if (rc) return rc;
  chol_subst(A, b, n, -1);
  return 0;
}

int la_chol_solve_banded(double *A, double *b, int n, int bw) {
  int rc = chol_factor(A, n, bw);
  if (rc) return rc;
  chol_subst(A, b, n, bw);
  return 0;
}

int la_lu_solve(double *A, double *b, int n) {
  for (int k = 0; k < n; ++k) {
    int p = k;
    double best = fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (fabs(A[i * n + k]) > best) {
        best = fabs(A[i * n + k]);
        p = i;
      }
    if (best == 0.0) return -1;
    if (p != k) {
      for (int j = 0; j < n; ++j) {
        double t = A[k * n + j];
        A[k * n + j] = A[p * n + j];
        A[p * n + j] = t;
      }
      double t = b[k];
      b[k] = b[p];
      b[p] = t;
    }
    for (int i = k + 1; i < n; ++i) {
      double l = A[i * n + k] / A[k * n + k];
      if (l != 0.0) {
        for (int j = k + 1; j < n; ++j) A[i * n + j] -= l * A[k * n + j];
        b[i] -= l * b[k];
      }
      A[i * n + k] = 0.0;
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int j = i + 1; j < n; ++j) s -= A[i * n + j] * b[j];
    b[i] = s / A[i * n + i];
  }
  return 0;
}

double la_norm2(const double *x, int n) {
  double s = 0.0;
  for (int i = 0; i < n; ++i) s += x[i] * x[i];
  return sqrt(s);
}
