// Stand-in for <OsqpEigen/OsqpEigen.h> (osqp-eigen and OSQP are absent from the image).
//
// TEST INFRASTRUCTURE ONLY: lets the reference's unmodified MheSrb.cpp / DecentralEst.cpp compile into oracle/_ref/.
// The Solver below accepts exactly the calls the reference makes (MheSrb.cpp:272-349, DecentralEst.cpp:204-217) and
// solves the QP   min 1/2 z'Hz + g'z   s.t.  l <= Az <= u   in one of two ways (refstub::osqp_mode()):
//   0 (default)  the exact optimum.  BASELINE.json defines parity against "both sides solving to eps 1e-8", i.e. the
//                unique optimum.  Rows with l == u are equalities, rows with |l|,|u| >= 1e20 are free (the reference's
//                placeholder +-OsqpEigen::INFTY, DecentralEst.cpp:474-481); the KKT system [[H A_e'],[A_e 0]] is
//                permuted to a band (variables and rows are both time ordered), factored by banded LU with partial
//                pivoting and polished by iterative refinement with a long-double residual.  A row with finite l < u
//                (never produced by the reference) aborts.
//   1            the OSQP-style ADMM of oracle/admm.c (restatement of the published algorithm, "parity unpinned")
//                with the settings handed to settings().
#pragma once
#include <Eigen/Sparse>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>

#include "../../oracle.h"

extern "C" int orc_admm_solve(int n, int m, const double *Hd, const double *gd, const double *Ad, const double *ld,
                              const double *ud, const orc_params *prm, double *z_out, int *iters_out);

namespace refstub {
inline int &osqp_mode() { static int m = 0; return m; }
inline int &osqp_last_iters() { static int m = 0; return m; }
inline int &osqp_last_bandwidth() { static int m = 0; return m; }
}  // namespace refstub

namespace OsqpEigen {

const double INFTY = 1e30;  // OSQP_INFTY
enum class ErrorExitFlag { NoError = 0, DataValidationError, SettingsValidationError, LinsysSolverLoadError,
                           LinsysSolverInitError, NonCvxError, MemAllocError, WorkspaceNotInitError };
enum class Status { Solved = 1, SolvedInaccurate = 2, MaxIterReached = -2 };

class Settings {
 public:
  void setWarmStart(bool v) { warm_start = v; }
  void setAdaptiveRho(bool v) { adaptive_rho = v; }
  void setVerbosity(bool v) { verbose = v; }
  void setPolish(bool v) { polish = v; }
  void setMaxIteration(int v) { max_iter = v; }
  void setRho(double v) { rho = v; }
  void setAlpha(double v) { alpha = v; }
  void setDelta(double v) { delta = v; }
  void setSigma(double v) { sigma = v; }
  void setRelativeTolerance(double v) { eps_rel = v; }
  void setAbsoluteTolerance(double v) { eps_abs = v; }
  void setPrimalInfeasibilityTolerance(double v) { eps_prim_inf = v; }
  void setDualInfeasibilityTolerance(double v) { eps_dual_inf = v; }
  void setPrimalInfeasibilityTollerance(double v) { eps_prim_inf = v; }
  void setDualInfeasibilityTollerance(double v) { eps_dual_inf = v; }
  void setTimeLimit(double v) { time_limit = v; }
  void setScaling(int) {}
  void setCheckTermination(int) {}
  bool warm_start = true, adaptive_rho = true, verbose = false, polish = false;
  int max_iter = 4000;
  double rho = 0.1, alpha = 1.6, delta = 1e-6, sigma = 1e-6, eps_rel = 1e-3, eps_abs = 1e-3, eps_prim_inf = 1e-4,
         eps_dual_inf = 1e-4, time_limit = 0.0;
};

class Data {
 public:
  void setNumberOfVariables(int n) { nV = n; }
  void setNumberOfConstraints(int m) { nC = m; }
  void clearHessianMatrix() { H = Eigen::Mat(); }
  void clearLinearConstraintsMatrix() { A = Eigen::Mat(); }
  bool setHessianMatrix(const Eigen::Mat &h) { H = h; return h.rows() == nV && h.cols() == nV; }
  bool setGradient(Eigen::Mat &gg) { g = gg; return gg.size() == nV; }
  bool setLinearConstraintsMatrix(const Eigen::Mat &a) { A = a; return a.rows() == nC && a.cols() == nV; }
  bool setLowerBound(Eigen::Mat &v) { l = v; return v.size() == nC; }
  bool setUpperBound(Eigen::Mat &v) { u = v; return v.size() == nC; }
  int nV = 0, nC = 0;
  Eigen::Mat H, A, g, l, u;
};

class Solver {
 public:
  Solver() : settings_(new Settings), data_(new Data) {}
  const std::unique_ptr<Settings> &settings() const { return settings_; }
  const std::unique_ptr<Data> &data() const { return data_; }
  void clearSolver() { init_ = false; }
  bool isInitialized() const { return init_; }
  bool initSolver() {
    init_ = data_->H.rows() == data_->nV && data_->A.rows() == data_->nC && data_->A.cols() == data_->nV &&
            data_->g.size() == data_->nV && data_->l.size() == data_->nC && data_->u.size() == data_->nC;
    return init_;
  }
  bool updateBounds(Eigen::Mat &lo, Eigen::Mat &up) { data_->l = lo; data_->u = up; return true; }
  bool updateGradient(Eigen::Mat &gg) { data_->g = gg; return true; }
  bool updateHessianMatrix(const Eigen::Mat &h) { data_->H = h; return true; }
  bool updateLinearConstraintsMatrix(const Eigen::Mat &a) { data_->A = a; return true; }
  const Eigen::VectorXd &getSolution() const { return sol_; }

  ErrorExitFlag solveProblem() {
    if (!init_) return ErrorExitFlag::WorkspaceNotInitError;
    if (refstub::osqp_mode() == 1) solve_admm(); else solve_exact();
    return ErrorExitFlag::NoError;
  }
  bool solve() { return solveProblem() == ErrorExitFlag::NoError; }

 private:
  void solve_admm() {
    const Data &d = *data_;
    int n = d.nV, m = d.nC;
    std::vector<double> Hd((size_t)n * n), Ad((size_t)m * n), z((size_t)n, 0.0);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Hd[(size_t)i * n + j] = d.H(i, j);
    for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) Ad[(size_t)i * n + j] = d.A(i, j);
    orc_params p;
    orc_params_go1_defaults(&p);
    p.rho = settings_->rho; p.alpha = settings_->alpha; p.delta = settings_->delta; p.sigma = settings_->sigma;
    p.adapt_rho = settings_->adaptive_rho; p.polish = settings_->polish; p.max_qp_iter = settings_->max_iter;
    p.relative_tol = settings_->eps_rel; p.abs_tol = settings_->eps_abs; p.prim_tol = settings_->eps_prim_inf;
    p.dual_tol = settings_->eps_dual_inf; p.time_limit = settings_->time_limit; p.verbose = 0;
    int iters = 0;
    orc_admm_solve(n, m, Hd.data(), d.g.data(), Ad.data(), d.l.data(), d.u.data(), &p, z.data(), &iters);
    refstub::osqp_last_iters() = iters;
    sol_ = Eigen::VectorXd(n);
    for (int i = 0; i < n; ++i) sol_(i) = z[(size_t)i];
  }

  void solve_exact() {
    const Data &d = *data_;
    const int n = d.nV;
    // equality rows; free rows are dropped
    std::vector<int> rows;
    for (int i = 0; i < d.nC; ++i) {
      double lo = d.l(i), up = d.u(i);
      if (lo <= -1e20 && up >= 1e20) continue;
      if (lo != up) { std::fprintf(stderr, "refstub OsqpEigen: row %d has l < u (%g, %g): not an equality\n", i, lo, up); std::abort(); }
      rows.push_back(i);
    }
    const int me = (int)rows.size(), K = n + me;
    // order: variables by index, each equality row right after the last variable it touches
    std::vector<std::pair<double, int>> key((size_t)K);
    for (int j = 0; j < n; ++j) key[(size_t)j] = {(double)j, j};
    for (int e = 0; e < me; ++e) {
      int last = 0;
      for (int j = 0; j < n; ++j) if (d.A(rows[(size_t)e], j) != 0.0) last = j;
      key[(size_t)(n + e)] = {(double)last + 0.5, n + e};
    }
    std::stable_sort(key.begin(), key.end(), [](const std::pair<double, int> &a, const std::pair<double, int> &b) { return a.first < b.first; });
    std::vector<int> pos((size_t)K);
    for (int k = 0; k < K; ++k) pos[(size_t)key[(size_t)k].second] = k;
    // dense storage of the permuted KKT matrix (row-major), bandwidth measured
    std::vector<double> M((size_t)K * K, 0.0), rhs((size_t)K, 0.0);
    int bw = 0;
    auto put = [&](int a, int b, double v) { M[(size_t)a * K + b] = v; int w = a > b ? a - b : b - a; if (w > bw) bw = w; };
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) { double v = d.H(i, j); if (v != 0.0) put(pos[(size_t)i], pos[(size_t)j], v); }
    for (int e = 0; e < me; ++e)
      for (int j = 0; j < n; ++j) {
        double v = d.A(rows[(size_t)e], j);
        if (v != 0.0) { put(pos[(size_t)(n + e)], pos[(size_t)j], v); put(pos[(size_t)j], pos[(size_t)(n + e)], v); }
      }
    for (int j = 0; j < n; ++j) rhs[(size_t)pos[(size_t)j]] = -d.g(j);
    for (int e = 0; e < me; ++e) rhs[(size_t)pos[(size_t)(n + e)]] = d.l(rows[(size_t)e]);
    refstub::osqp_last_bandwidth() = bw;
    std::vector<double> M0(M), x((size_t)K, 0.0);
    // banded LU with partial pivoting (kl = ku = bw; fill-in to ku + kl above the diagonal)
    std::vector<int> piv((size_t)K);
    for (int k = 0; k < K; ++k) {
      int iend = std::min(K - 1, k + bw), jend = std::min(K - 1, k + 2 * bw);
      int p = k;
      double best = std::fabs(M[(size_t)k * K + k]);
      for (int i = k + 1; i <= iend; ++i) { double a = std::fabs(M[(size_t)i * K + k]); if (a > best) { best = a; p = i; } }
      piv[(size_t)k] = p;
      if (best == 0.0) { std::fprintf(stderr, "refstub OsqpEigen: singular KKT at %d\n", k); std::abort(); }
      if (p != k) for (int j = k; j <= jend; ++j) std::swap(M[(size_t)k * K + j], M[(size_t)p * K + j]);
      double dk = M[(size_t)k * K + k];
      for (int i = k + 1; i <= iend; ++i) {
        double f = M[(size_t)i * K + k] / dk;
        M[(size_t)i * K + k] = f;
        if (f != 0.0) for (int j = k + 1; j <= jend; ++j) M[(size_t)i * K + j] -= f * M[(size_t)k * K + j];
      }
    }
    auto lu_solve = [&](std::vector<double> &b) {
      for (int k = 0; k < K; ++k) {
        int p = piv[(size_t)k];
        if (p != k) std::swap(b[(size_t)k], b[(size_t)p]);
        int iend = std::min(K - 1, k + bw);
        for (int i = k + 1; i <= iend; ++i) b[(size_t)i] -= M[(size_t)i * K + k] * b[(size_t)k];
      }
      for (int k = K - 1; k >= 0; --k) {
        int jend = std::min(K - 1, k + 2 * bw);
        double s = b[(size_t)k];
        for (int j = k + 1; j <= jend; ++j) s -= M[(size_t)k * K + j] * b[(size_t)j];
        b[(size_t)k] = s / M[(size_t)k * K + k];
      }
    };
    x = rhs;
    lu_solve(x);
    for (int it = 0; it < 3; ++it) {  // iterative refinement, residual in long double
      std::vector<double> r((size_t)K);
      for (int a = 0; a < K; ++a) {
        long double s = rhs[(size_t)a];
        int j0 = std::max(0, a - bw), j1 = std::min(K - 1, a + bw);
        for (int j = j0; j <= j1; ++j) s -= (long double)M0[(size_t)a * K + j] * (long double)x[(size_t)j];
        r[(size_t)a] = (double)s;
      }
      lu_solve(r);
      for (int a = 0; a < K; ++a) x[(size_t)a] += r[(size_t)a];
    }
    sol_ = Eigen::VectorXd(n);
    for (int j = 0; j < n; ++j) sol_(j) = x[(size_t)pos[(size_t)j]];
  }

  std::unique_ptr<Settings> settings_;
  std::unique_ptr<Data> data_;
  Eigen::VectorXd sol_;
  bool init_ = false;
};

}  // namespace OsqpEigen
