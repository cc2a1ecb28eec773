/* Dense row-major helpers for the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle/README.md):
 * nothing under oracle/ is linked, imported or executed by the product path. */
#ifndef ORC_LA_H
#define ORC_LA_H
#ifdef __cplusplus
extern "C" {
#endif

/* C[m x n] = A[m x k] * B[k x n] */
void la_mm(double *C, const double *A, const double *B, int m, int k, int n);
/* C[m x n] = A[m x k] * B^T, B is [n x k] */
void la_mmt(double *C, const double *A, const double *B, int m, int k, int n);
/* C[m x n] = A^T * B, A is [k x m], B is [k x n] */
void la_mtm(double *C, const double *A, const double *B, int m, int k, int n);
void la_transpose(double *At, const double *A, int m, int n);
void la_zero(double *A, int n);
void la_eye(double *A, int n);
void la_copy(double *dst, const double *src, int n);
/* dst[r0+i][c0+j] = src[i][j] */
void la_set_block(double *dst, int ld, int r0, int c0, const double *src, int m, int n);
void la_get_block(double *dst, const double *src, int ld, int r0, int c0, int m, int n);
/* General inverse by LU with partial pivoting (what Eigen's MatrixXd::inverse() does for dynamic
 * sizes, PartialPivLU).  Returns 0 on success, -1 if a zero pivot is met. */
int la_inverse(double *Ainv, const double *A, int n);
/* SPD inverse through Cholesky (Eigen SimplicialLLT::solve(I) equivalent). Returns 0 / -1. */
int la_spd_inverse(double *Ainv, const double *A, int n);
/* In-place banded-unaware dense Cholesky solve of A x = b (A SPD, destroyed). Returns 0 / -1. */
int la_chol_solve(double *A, double *b, int n);
/* Symmetric banded Cholesky solve: A dense storage n x n but only |i-j|<=bw touched. */
int la_chol_solve_banded(double *A, double *b, int n, int bw);
/* LU solve with partial pivoting of A x = b (A destroyed, b overwritten). Returns 0 / -1. */
int la_lu_solve(double *A, double *b, int n);
double la_norm2(const double *x, int n);

#ifdef __cplusplus
}
#endif
#endif
