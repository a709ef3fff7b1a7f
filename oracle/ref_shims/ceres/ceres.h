// ceres/ceres.h — SHIM ("mini-Ceres"; test infrastructure written for this repo, NOT Ceres Solver).
//
// The reference's hot path (multi-view-refinement/cost.cc, solve.cc) is written against Ceres
// Solver, which is neither vendored by the reference nor present in this image.  So that the
// reference's own sources can be compiled UNMODIFIED, from where they lie under /root/reference
// (oracle/build_ref.py), this header declares the slice of the Ceres 1.14 API those files use:
//
//   cost.cc:56-63   ceres::Jet members .a / .v
//   cost.cc:92-94   ceres::CostFunction, ceres::AutoDiffCostFunction<F, 2, 2, 2>
//   solve.cc:92     ceres::Problem
//   solve.cc:107-122  Problem::AddResidualBlock, ScaledLoss, CauchyLoss, TukeyLoss, TAKE_OWNERSHIP
//   solve.cc:134-140  SetParameterBlockConstant, SetParameter{Lower,Upper}Bound
//   solve.cc:146-159  Solver::Options / Summary, ceres::Solve, SPARSE_NORMAL_CHOLESKY
//
// The behaviour behind the declarations (oracle/ref_shims/mini_ceres.cc) is a restatement, from
// the published algorithm of Ceres Solver 1.14, of: automatic differentiation with Jets, the
// loss functions and the Corrector, program reduction (constant blocks, fixed cost), the
// bounds-projecting Plus, TrustRegionMinimizer + LevenbergMarquardtStrategy +
// TrustRegionStepEvaluator, ArmijoLineSearch with CUBIC interpolation, and polynomial.cc
// (Vandermonde fit by full-pivot LU, roots by balanced companion matrix eigenvalues).  It is the
// same specification as SURVEY.md Appendix A, written literally and generically (it knows
// nothing about flow grids).  What it does NOT reproduce bit for bit: Eigen's and
// SuiteSparse's floating-point summation orders (AMD parameter ordering, supernodal Cholesky)
// — round-off level only.
#ifndef LFR_SHIM_CERES_H_
#define LFR_SHIM_CERES_H_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <limits>
#include <map>
#include <string>
#include <vector>

namespace ceres {

enum Ownership { DO_NOT_TAKE_OWNERSHIP, TAKE_OWNERSHIP };
enum LinearSolverType {
  DENSE_NORMAL_CHOLESKY,
  DENSE_QR,
  SPARSE_NORMAL_CHOLESKY,
  DENSE_SCHUR,
  SPARSE_SCHUR,
  ITERATIVE_SCHUR,
  CGNR
};
enum TerminationType { CONVERGENCE, NO_CONVERGENCE, FAILURE, USER_SUCCESS, USER_FAILURE };

// ---- Jet (jet.h): a + sum_i v[i] e_i -------------------------------------------------------
template <typename T, int N>
struct JetVec {
  T d[N];
  JetVec() {
    for (int i = 0; i < N; ++i) d[i] = T(0);
  }
  T& operator[](int i) { return d[i]; }
  const T& operator[](int i) const { return d[i]; }
};
template <typename T, int N>
inline JetVec<T, N> operator*(const T& s, const JetVec<T, N>& v) {
  JetVec<T, N> r;
  for (int i = 0; i < N; ++i) r.d[i] = s * v.d[i];
  return r;
}
template <typename T, int N>
inline JetVec<T, N> operator*(const JetVec<T, N>& v, const T& s) {
  JetVec<T, N> r;
  for (int i = 0; i < N; ++i) r.d[i] = v.d[i] * s;
  return r;
}
template <typename T, int N>
inline JetVec<T, N> operator+(const JetVec<T, N>& a, const JetVec<T, N>& b) {
  JetVec<T, N> r;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template <typename T, int N>
inline JetVec<T, N> operator-(const JetVec<T, N>& a, const JetVec<T, N>& b) {
  JetVec<T, N> r;
  for (int i = 0; i < N; ++i) r.d[i] = a.d[i] - b.d[i];
  return r;
}

template <typename T, int N>
struct Jet {
  enum { DIMENSION = N };
  T a;
  JetVec<T, N> v;
  Jet() : a() {}
  explicit Jet(const T& value) : a(value) {}
  Jet(const T& value, int k) : a(value) { v[k] = T(1); }
};
template <typename T, int N>
inline Jet<T, N> operator-(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a - g.a;
  h.v = f.v - g.v;
  return h;
}
template <typename T, int N>
inline Jet<T, N> operator+(const Jet<T, N>& f, const Jet<T, N>& g) {
  Jet<T, N> h;
  h.a = f.a + g.a;
  h.v = f.v + g.v;
  return h;
}

// ---- CostFunction / AutoDiffCostFunction ------------------------------------------------------
class CostFunction {
 public:
  CostFunction() : num_residuals_(0) {}
  virtual ~CostFunction() {}
  // jacobians[i] (may be NULL) is num_residuals x block_size(i), row-major
  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
  const std::vector<int>& parameter_block_sizes() const { return parameter_block_sizes_; }
  int num_residuals() const { return num_residuals_; }

 protected:
  std::vector<int>* mutable_parameter_block_sizes() { return &parameter_block_sizes_; }
  void set_num_residuals(int n) { num_residuals_ = n; }

 private:
  std::vector<int> parameter_block_sizes_;
  int num_residuals_;
};

// Two parameter blocks are all the reference uses (cost.cc:93).
template <typename CostFunctor, int kNumResiduals, int N0, int N1>
class AutoDiffCostFunction : public CostFunction {
 public:
  explicit AutoDiffCostFunction(CostFunctor* functor) : functor_(functor) {
    set_num_residuals(kNumResiduals);
    mutable_parameter_block_sizes()->push_back(N0);
    mutable_parameter_block_sizes()->push_back(N1);
  }
  virtual ~AutoDiffCostFunction() { delete functor_; }

  virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const {
    if (!jacobians) return (*functor_)(parameters[0], parameters[1], residuals);
    typedef Jet<double, N0 + N1> JetT;
    JetT x0[N0], x1[N1], out[kNumResiduals];
    for (int i = 0; i < N0; ++i) x0[i] = JetT(parameters[0][i], i);
    for (int i = 0; i < N1; ++i) x1[i] = JetT(parameters[1][i], N0 + i);
    if (!(*functor_)(x0, x1, out)) return false;
    for (int r = 0; r < kNumResiduals; ++r) {
      residuals[r] = out[r].a;
      if (jacobians[0])
        for (int c = 0; c < N0; ++c) jacobians[0][r * N0 + c] = out[r].v[c];
      if (jacobians[1])
        for (int c = 0; c < N1; ++c) jacobians[1][r * N1 + c] = out[r].v[N0 + c];
    }
    return true;
  }

 private:
  CostFunctor* functor_;
};

// ---- loss functions (loss_function.h / .cc) ----------------------------------------------------
class LossFunction {
 public:
  virtual ~LossFunction() {}
  virtual void Evaluate(double sq_norm, double out[3]) const = 0;
};

class CauchyLoss : public LossFunction {
 public:
  explicit CauchyLoss(double a) : b_(a * a), c_(1 / b_) {}
  virtual void Evaluate(double, double*) const;

 private:
  const double b_, c_;
};

class TukeyLoss : public LossFunction {
 public:
  explicit TukeyLoss(double a) : a_squared_(a * a) {}
  virtual void Evaluate(double, double*) const;

 private:
  const double a_squared_;
};

class ScaledLoss : public LossFunction {
 public:
  ScaledLoss(const LossFunction* rho, double a, Ownership ownership) : rho_(rho), a_(a), ownership_(ownership) {}
  virtual ~ScaledLoss() {
    if (ownership_ == TAKE_OWNERSHIP) delete rho_;
  }
  virtual void Evaluate(double, double*) const;

 private:
  const LossFunction* rho_;
  const double a_;
  const Ownership ownership_;
};

// ---- Problem ------------------------------------------------------------------------------------
namespace internal {
struct ProblemImpl;
}

class Problem {
 public:
  Problem();
  ~Problem();
  void* AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, double* x0, double* x1);
  void SetParameterBlockConstant(double* values);
  void SetParameterLowerBound(double* values, int index, double lower_bound);
  void SetParameterUpperBound(double* values, int index, double upper_bound);
  internal::ProblemImpl* impl() { return impl_; }

 private:
  Problem(const Problem&);
  void operator=(const Problem&);
  internal::ProblemImpl* impl_;
};

// ---- Solver -------------------------------------------------------------------------------------
class Solver {
 public:
  struct Options {
    Options()
        : max_num_iterations(50),
          num_threads(1),
          linear_solver_type(SPARSE_NORMAL_CHOLESKY),
          minimizer_progress_to_stdout(false),
          max_num_consecutive_invalid_steps(5),
          function_tolerance(1e-6),
          gradient_tolerance(1e-10),
          parameter_tolerance(1e-8),
          initial_trust_region_radius(1e4),
          max_trust_region_radius(1e16),
          min_trust_region_radius(1e-32),
          min_relative_decrease(1e-3),
          min_lm_diagonal(1e-6),
          max_lm_diagonal(1e32),
          jacobi_scaling(true),
          line_search_sufficient_function_decrease(1e-4),
          max_line_search_step_contraction(1e-3),
          min_line_search_step_contraction(0.6),
          max_num_line_search_step_size_iterations(20),
          min_line_search_step_size(1e-9) {}
    int max_num_iterations;
    int num_threads;
    LinearSolverType linear_solver_type;
    bool minimizer_progress_to_stdout;
    int max_num_consecutive_invalid_steps;
    double function_tolerance, gradient_tolerance, parameter_tolerance;
    double initial_trust_region_radius, max_trust_region_radius, min_trust_region_radius;
    double min_relative_decrease, min_lm_diagonal, max_lm_diagonal;
    bool jacobi_scaling;
    double line_search_sufficient_function_decrease;
    double max_line_search_step_contraction, min_line_search_step_contraction;
    int max_num_line_search_step_size_iterations;
    double min_line_search_step_size;
  };
  struct Summary {
    Summary()
        : termination_type(FAILURE),
          initial_cost(-1),
          final_cost(-1),
          fixed_cost(-1),
          num_iterations(0),
          num_successful_steps(0),
          num_line_search_steps(0),
          num_parameter_blocks_reduced(0),
          num_residual_blocks_reduced(0) {}
    bool IsSolutionUsable() const {
      return termination_type == CONVERGENCE || termination_type == NO_CONVERGENCE ||
             termination_type == USER_SUCCESS;
    }
    TerminationType termination_type;
    std::string message;
    double initial_cost, final_cost, fixed_cost;
    int num_iterations;  // iterations run by the minimizer (= iterations.size() - 1 in Ceres)
    int num_successful_steps, num_line_search_steps;
    int num_parameter_blocks_reduced, num_residual_blocks_reduced;
  };
};

void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);

// ---- knobs and probes of the shim (not Ceres API) ----------------------------------------------
namespace shim {
// TukeyLoss changed between Ceres 1.x (rho = a^2/6 (1 - v^3)) and 2.x (a^2/3 (1 - v^3)).
// 1 (default) or 2; also read from the environment variable LFR_CERES_TUKEY_VARIANT at start-up.
void SetTukeyVariant(int variant);
// Every ceres::Solve() appends its summary here when recording is on (solve.cc discards it).
struct Record {
  const double* first_parameter_block;  // identifies the problem: address of the first block added
  Solver::Summary summary;
};
void SetRecording(bool on);
std::vector<Record> TakeRecords();
// polynomial.cc, exposed for the tests: samples = {x, value, gradient, value_valid, gradient_valid}
void MinimizeInterpolatingPolynomial(const double* samples5, int n_samples, double x_min, double x_max,
                                     double* optimal_x, double* optimal_value);
// real parts of all roots (FindPolynomialRoots), coefficients highest degree first; returns count or -1
int FindPolynomialRoots(const double* poly, int n_coeff, double* real, double* imag);
}  // namespace shim

}  // namespace ceres

#endif  // LFR_SHIM_CERES_H_
