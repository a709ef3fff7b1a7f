// mini_ceres.cc — behaviour behind oracle/ref_shims/ceres/ceres.h.
//
// TEST INFRASTRUCTURE ONLY (see the header): a literal, generic restatement of the slice of
// Ceres Solver 1.14 that multi-view-refinement/solve.cc:79-160 exercises.  Nothing under
// local-feature-refinement_b200/ may link this.  Section names below are the Ceres source files
// whose published algorithm each block restates (none of them is in /root/reference: Ceres is an
// un-vendored dependency, CMakeLists.txt:9).
#include "ceres/ceres.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

namespace ceres {

// ================================ loss_function.cc ================================
namespace {
int g_tukey_variant = 0;  // 0 = not initialised
int tukey_variant() {
  if (g_tukey_variant == 0) {
    const char* e = std::getenv("LFR_CERES_TUKEY_VARIANT");
    g_tukey_variant = (e && std::atoi(e) == 2) ? 2 : 1;
  }
  return g_tukey_variant;
}
}  // namespace

void CauchyLoss::Evaluate(double s, double rho[3]) const {
  const double sum = 1.0 + s * c_;
  const double inv = 1.0 / sum;
  // 'sum' and 'inv' are always positive, assuming that 's' is.
  rho[0] = b_ * std::log(sum);
  rho[1] = std::max(std::numeric_limits<double>::min(), inv);
  rho[2] = -c_ * (inv * inv);
}

void TukeyLoss::Evaluate(double s, double* rho) const {
  const bool v2 = tukey_variant() == 2;
  if (s <= a_squared_) {  // inlier region
    const double value = 1.0 - s / a_squared_;
    const double value_sq = value * value;
    if (v2) {  // Ceres 2.x
      rho[0] = a_squared_ / 3.0 * (1.0 - value_sq * value);
      rho[1] = value_sq;
      rho[2] = -2.0 / a_squared_ * value;
    } else {  // Ceres 1.x
      rho[0] = a_squared_ / 6.0 * (1.0 - value_sq * value);
      rho[1] = 0.5 * value_sq;
      rho[2] = -1.0 / a_squared_ * value;
    }
  } else {  // outlier region
    rho[0] = v2 ? a_squared_ / 3.0 : a_squared_ / 6.0;
    rho[1] = 0.0;
    rho[2] = 0.0;
  }
}

void ScaledLoss::Evaluate(double s, double rho[3]) const {
  if (rho_ == NULL) {
    rho[0] = a_ * s;
    rho[1] = a_;
    rho[2] = 0.0;
  } else {
    rho_->Evaluate(s, rho);
    rho[0] *= a_;
    rho[1] *= a_;
    rho[2] *= a_;
  }
}

// ================================ problem_impl.cc ================================
namespace internal {

struct ParameterBlock {
  double* user_state;
  int size;
  bool is_constant;
  std::vector<double> lower, upper;  // empty = unbounded
  int reduced_offset;                // offset in the reduced state vector, -1 if constant / unused
};

struct ResidualBlock {
  CostFunction* cost_function;
  LossFunction* loss_function;
  int block[2];  // indices into ProblemImpl::parameter_blocks
};

struct ProblemImpl {
  std::vector<ParameterBlock> parameter_blocks;  // in order of first appearance (Program order)
  std::map<double*, int> index_of;
  std::vector<ResidualBlock> residual_blocks;    // in order of AddResidualBlock

  int intern(double* values, int size) {
    std::map<double*, int>::iterator it = index_of.find(values);
    if (it != index_of.end()) return it->second;
    ParameterBlock pb;
    pb.user_state = values;
    pb.size = size;
    pb.is_constant = false;
    pb.reduced_offset = -1;
    parameter_blocks.push_back(pb);
    index_of[values] = (int)parameter_blocks.size() - 1;
    return (int)parameter_blocks.size() - 1;
  }
  int find(double* values) const {
    std::map<double*, int>::const_iterator it = index_of.find(values);
    if (it == index_of.end()) {
      std::fprintf(stderr, "mini-ceres: parameter block not found in the problem\n");
      std::abort();  // Ceres: LOG(FATAL)
    }
    return it->second;
  }
};

}  // namespace internal

Problem::Problem() : impl_(new internal::ProblemImpl) {}
Problem::~Problem() {
  // Problem::Options defaults: the problem owns cost and loss functions.
  for (size_t i = 0; i < impl_->residual_blocks.size(); ++i) {
    delete impl_->residual_blocks[i].cost_function;
    delete impl_->residual_blocks[i].loss_function;
  }
  delete impl_;
}

void* Problem::AddResidualBlock(CostFunction* cost_function, LossFunction* loss_function, double* x0, double* x1) {
  internal::ResidualBlock rb;
  rb.cost_function = cost_function;
  rb.loss_function = loss_function;
  rb.block[0] = impl_->intern(x0, cost_function->parameter_block_sizes()[0]);
  rb.block[1] = impl_->intern(x1, cost_function->parameter_block_sizes()[1]);
  impl_->residual_blocks.push_back(rb);
  return &impl_->residual_blocks.back();
}
void Problem::SetParameterBlockConstant(double* values) { impl_->parameter_blocks[impl_->find(values)].is_constant = true; }
void Problem::SetParameterLowerBound(double* values, int index, double lower_bound) {
  internal::ParameterBlock& pb = impl_->parameter_blocks[impl_->find(values)];
  if (pb.lower.empty()) pb.lower.assign(pb.size, -std::numeric_limits<double>::max());
  pb.lower[index] = lower_bound;
}
void Problem::SetParameterUpperBound(double* values, int index, double upper_bound) {
  internal::ParameterBlock& pb = impl_->parameter_blocks[impl_->find(values)];
  if (pb.upper.empty()) pb.upper.assign(pb.size, std::numeric_limits<double>::max());
  pb.upper[index] = upper_bound;
}

// ================================ polynomial.cc ================================
namespace {

typedef std::vector<double> Vec;

inline double EvaluatePolynomial(const Vec& polynomial, double x) {
  double v = 0.0;
  for (size_t i = 0; i < polynomial.size(); ++i) v = v * x + polynomial[i];
  return v;
}

Vec RemoveLeadingZeros(const Vec& polynomial_in) {
  size_t i = 0;
  while (i < (polynomial_in.size() - 1) && polynomial_in[i] == 0.0) ++i;
  return Vec(polynomial_in.begin() + i, polynomial_in.end());
}

Vec DifferentiatePolynomial(const Vec& polynomial) {
  const int degree = (int)polynomial.size() - 1;
  if (degree == 0) return Vec(1, 0.0);  // degree zero polynomials are constants: derivative 0
  Vec derivative(degree);
  for (int i = 0; i < degree; ++i) derivative[i] = (degree - i) * polynomial[i];
  return derivative;
}

void FindLinearPolynomialRoots(const Vec& polynomial, Vec* real, Vec* imag) {
  real->assign(1, -polynomial[1] / polynomial[0]);
  imag->assign(1, 0.0);
}

void FindQuadraticPolynomialRoots(const Vec& polynomial, Vec* real, Vec* imag) {
  const double a = polynomial[0], b = polynomial[1], c = polynomial[2];
  const double D = b * b - 4 * a * c;
  const double sqrt_D = std::sqrt(std::fabs(D));
  real->assign(2, 0.0);
  imag->assign(2, 0.0);
  // Real roots.
  if (D >= 0) {
    // Stable quadratic roots according to BKP Horn.
    // http://people.csail.mit.edu/bkph/articles/Quadratics.pdf
    if (b >= 0) {
      (*real)[0] = (-b - sqrt_D) / (2.0 * a);
      (*real)[1] = (2.0 * c) / (-b - sqrt_D);
    } else {
      (*real)[0] = (2.0 * c) / (-b + sqrt_D);
      (*real)[1] = (-b + sqrt_D) / (2.0 * a);
    }
    return;
  }
  // Use the normal quadratic formula for the complex case.
  (*real)[0] = -b / (2.0 * a);
  (*real)[1] = -b / (2.0 * a);
  (*imag)[0] = sqrt_D / (2.0 * a);
  (*imag)[1] = -sqrt_D / (2.0 * a);
}

// Balancing of the companion matrix (Parlett & Reinsch's norm-reducing diagonal similarity with
// radix-2 scale factors and gamma = 0.9), on the off-diagonal part.  m is n x n row-major.
void BalanceCompanionMatrix(std::vector<double>* m_ptr, int n) {
  std::vector<double>& m = *m_ptr;
  std::vector<double> off(m);
  for (int i = 0; i < n; ++i) off[i * n + i] = 0.0;
  const double gamma = 0.9;
  bool scaling_has_changed;
  do {
    scaling_has_changed = false;
    for (int i = 0; i < n; ++i) {
      double row_norm = 0.0, col_norm = 0.0;
      for (int j = 0; j < n; ++j) {
        row_norm += std::fabs(off[i * n + j]);
        col_norm += std::fabs(off[j * n + i]);
      }
      // decompose row_norm / col_norm into mantissa * 2^exponent, 0.5 <= mantissa < 1; only the exponent is used
      int exponent = 0;
      std::frexp(row_norm / col_norm, &exponent);
      exponent /= 2;
      if (exponent != 0) {
        const double scaled_col_norm = std::ldexp(col_norm, exponent);
        const double scaled_row_norm = std::ldexp(row_norm, -exponent);
        if (scaled_col_norm + scaled_row_norm < gamma * (col_norm + row_norm)) {
          // accept the new scaling (powers of two: exact)
          scaling_has_changed = true;
          for (int j = 0; j < n; ++j) {
            off[i * n + j] *= std::ldexp(1.0, -exponent);
            off[j * n + i] *= std::ldexp(1.0, exponent);
          }
        }
      }
    }
  } while (scaling_has_changed);
  for (int i = 0; i < n; ++i) off[i * n + i] = m[i * n + i];
  m = off;
}

// Eigenvalues of a real upper-Hessenberg matrix by the shifted (Francis double step) QR
// iteration — the algorithm behind Eigen::EigenSolver / RealSchur (EISPACK hqr).  a is n x n
// row-major and is destroyed.  Eigen's exact shift/deflation details differ at round-off level.
bool HessenbergEigenvalues(std::vector<double>& A, int n, Vec* wr_out, Vec* wi_out) {
  Vec& wr = *wr_out;
  Vec& wi = *wi_out;
  wr.assign(n, 0.0);
  wi.assign(n, 0.0);
#define a_(i, j) A[(i) * n + (j)]
  int nn, m, l, k, j, its, i, mmin;
  double z = 0, y, x, w, v, u, t, s, r = 0, q = 0, p = 0, anorm = 0.0;
  for (i = 0; i < n; i++)
    for (j = std::max(i - 1, 0); j < n; j++) anorm += std::fabs(a_(i, j));
  nn = n - 1;
  t = 0.0;
  while (nn >= 0) {
    its = 0;
    do {
      for (l = nn; l >= 1; l--) {  // look for a single small subdiagonal element
        s = std::fabs(a_(l - 1, l - 1)) + std::fabs(a_(l, l));
        if (s == 0.0) s = anorm;
        if (std::fabs(a_(l, l - 1)) + s == s) {
          a_(l, l - 1) = 0.0;
          break;
        }
      }
      x = a_(nn, nn);
      if (l == nn) {  // one root found
        wr[nn] = x + t;
        wi[nn--] = 0.0;
      } else {
        y = a_(nn - 1, nn - 1);
        w = a_(nn, nn - 1) * a_(nn - 1, nn);
        if (l == nn - 1) {  // two roots found
          p = 0.5 * (y - x);
          q = p * p + w;
          z = std::sqrt(std::fabs(q));
          x += t;
          if (q >= 0.0) {  // a real pair
            z = p + (p >= 0.0 ? std::fabs(z) : -std::fabs(z));
            wr[nn - 1] = wr[nn] = x + z;
            if (z != 0.0) wr[nn] = x - w / z;
            wi[nn - 1] = wi[nn] = 0.0;
          } else {  // a complex pair
            wr[nn - 1] = wr[nn] = x + p;
            wi[nn - 1] = -(wi[nn] = z);
          }
          nn -= 2;
        } else {  // no roots found: continue the iteration
          if (its == 60) return false;
          if (its == 10 || its == 20) {  // exceptional shift
            t += x;
            for (i = 0; i <= nn; i++) a_(i, i) -= x;
            s = std::fabs(a_(nn, nn - 1)) + std::fabs(a_(nn - 1, nn - 2));
            y = x = 0.75 * s;
            w = -0.4375 * s * s;
          }
          ++its;
          for (m = nn - 2; m >= l; m--) {  // form the shift, look for two consecutive small subdiagonal elements
            z = a_(m, m);
            r = x - z;
            s = y - z;
            p = (r * s - w) / a_(m + 1, m) + a_(m, m + 1);
            q = a_(m + 1, m + 1) - z - r - s;
            r = a_(m + 2, m + 1);
            s = std::fabs(p) + std::fabs(q) + std::fabs(r);
            p /= s;
            q /= s;
            r /= s;
            if (m == l) break;
            u = std::fabs(a_(m, m - 1)) * (std::fabs(q) + std::fabs(r));
            v = std::fabs(p) * (std::fabs(a_(m - 1, m - 1)) + std::fabs(z) + std::fabs(a_(m + 1, m + 1)));
            if (u + v == v) break;
          }
          for (i = m + 2; i <= nn; i++) {
            a_(i, i - 2) = 0.0;
            if (i != m + 2) a_(i, i - 3) = 0.0;
          }
          for (k = m; k <= nn - 1; k++) {  // double QR step on rows l..nn and columns m..nn
            if (k != m) {
              p = a_(k, k - 1);
              q = a_(k + 1, k - 1);
              r = 0.0;
              if (k != nn - 1) r = a_(k + 2, k - 1);
              if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0) {
                p /= x;
                q /= x;
                r /= x;
              }
            }
            const double nrm = std::sqrt(p * p + q * q + r * r);
            s = (p >= 0.0) ? nrm : -nrm;
            if (s != 0.0) {
              if (k == m) {
                if (l != m) a_(k, k - 1) = -a_(k, k - 1);
              } else {
                a_(k, k - 1) = -s * x;
              }
              p += s;
              x = p / s;
              y = q / s;
              z = r / s;
              q /= p;
              r /= p;
              for (j = k; j <= nn; j++) {  // row modification
                p = a_(k, j) + q * a_(k + 1, j);
                if (k != nn - 1) {
                  p += r * a_(k + 2, j);
                  a_(k + 2, j) -= p * z;
                }
                a_(k + 1, j) -= p * y;
                a_(k, j) -= p * x;
              }
              mmin = nn < k + 3 ? nn : k + 3;
              for (i = l; i <= mmin; i++) {  // column modification
                p = x * a_(i, k) + y * a_(i, k + 1);
                if (k != nn - 1) {
                  p += z * a_(i, k + 2);
                  a_(i, k + 2) -= p * r;
                }
                a_(i, k + 1) -= p * q;
                a_(i, k) -= p;
              }
            }
          }
        }
      }
    } while (l < nn - 1);
  }
#undef a_
  return true;
}

bool FindPolynomialRootsImpl(const Vec& polynomial_in, Vec* real, Vec* imag) {
  if (polynomial_in.empty()) return false;  // "Invalid polynomial of size 0 passed to FindPolynomialRoots"
  Vec polynomial = RemoveLeadingZeros(polynomial_in);
  const int degree = (int)polynomial.size() - 1;
  real->clear();
  imag->clear();
  // Is the polynomial constant?
  if (degree == 0) return true;  // "Trying to extract roots from a constant polynomial"
  // Linear
  if (degree == 1) {
    FindLinearPolynomialRoots(polynomial, real, imag);
    return true;
  }
  // Quadratic
  if (degree == 2) {
    FindQuadraticPolynomialRoots(polynomial, real, imag);
    return true;
  }
  // The degree is now known to be at least 3.  Divide by the leading term, build the companion
  // matrix (ones on the subdiagonal, last column = minus the reversed coefficients), balance it,
  // and take its eigenvalues.
  const double leading_term = polynomial[0];
  for (size_t i = 0; i < polynomial.size(); ++i) polynomial[i] /= leading_term;
  std::vector<double> companion((size_t)degree * degree, 0.0);
  for (int i = 1; i < degree; ++i) companion[i * degree + (i - 1)] = 1.0;
  for (int i = 0; i < degree; ++i) companion[i * degree + (degree - 1)] = -polynomial[degree - i];
  BalanceCompanionMatrix(&companion, degree);
  if (!HessenbergEigenvalues(companion, degree, real, imag)) return false;  // "Failed to extract eigenvalues from companion matrix."
  return true;
}

void MinimizePolynomial(const Vec& polynomial, double x_min, double x_max, double* optimal_x, double* optimal_value) {
  // Find the minimum of the polynomial at the two ends.  We start by inspecting the middle of
  // the interval (technically not needed; keeps the code close to the minFunc package).
  *optimal_x = (x_min + x_max) / 2.0;
  *optimal_value = EvaluatePolynomial(polynomial, *optimal_x);
  const double x_min_value = EvaluatePolynomial(polynomial, x_min);
  if (x_min_value < *optimal_value) {
    *optimal_value = x_min_value;
    *optimal_x = x_min;
  }
  const double x_max_value = EvaluatePolynomial(polynomial, x_max);
  if (x_max_value < *optimal_value) {
    *optimal_value = x_max_value;
    *optimal_x = x_max;
  }
  // If the polynomial is linear or constant, we are done.
  if (polynomial.size() <= 2) return;
  const Vec derivative = DifferentiatePolynomial(polynomial);
  Vec roots_real, roots_imag;
  if (!FindPolynomialRootsImpl(derivative, &roots_real, &roots_imag)) return;  // LOG(WARNING): unable to find the critical points
  // real parts of ALL roots are candidates
  for (size_t i = 0; i < roots_real.size(); ++i) {
    const double root = roots_real[i];
    if ((root < x_min) || (root > x_max)) continue;
    const double value = EvaluatePolynomial(polynomial, root);
    if (value < *optimal_value) {
      *optimal_value = value;
      *optimal_x = root;
    }
  }
}

struct FunctionSample {
  FunctionSample() : x(0.0), value(0.0), value_is_valid(false), gradient(0.0), gradient_is_valid(false) {}
  double x, value;
  bool value_is_valid;
  double gradient;
  bool gradient_is_valid;
};

// lhs.fullPivLu().solve(rhs): Gaussian elimination with complete pivoting, rank decided by
// |pivot| > eps * n * max|pivot| as in Eigen's FullPivLU, free variables set to zero.
Vec FullPivLuSolve(std::vector<double> lu, int n, const Vec& rhs) {
  std::vector<int> row_t(n), col_t(n);
  int nonzero_pivots = n;
  double maxpivot = 0.0;
  for (int k = 0; k < n; ++k) {
    // biggest coefficient of the remaining corner; Eigen scans column by column, first maximum wins
    int br = k, bc = k;
    double biggest = -1.0;
    for (int c = k; c < n; ++c)
      for (int r = k; r < n; ++r) {
        const double v = std::fabs(lu[r * n + c]);
        if (v > biggest) {
          biggest = v;
          br = r;
          bc = c;
        }
      }
    if (biggest == 0.0) {
      nonzero_pivots = k;
      for (int i = k; i < n; ++i) {
        row_t[i] = i;
        col_t[i] = i;
      }
      break;
    }
    if (biggest > maxpivot) maxpivot = biggest;
    row_t[k] = br;
    col_t[k] = bc;
    if (k != br)
      for (int c = 0; c < n; ++c) std::swap(lu[k * n + c], lu[br * n + c]);
    if (k != bc)
      for (int r = 0; r < n; ++r) std::swap(lu[r * n + k], lu[r * n + bc]);
    if (k < n - 1) {
      for (int r = k + 1; r < n; ++r) lu[r * n + k] /= lu[k * n + k];
      for (int r = k + 1; r < n; ++r)
        for (int c = k + 1; c < n; ++c) lu[r * n + c] -= lu[r * n + k] * lu[k * n + c];
    }
  }
  // rank
  const double threshold = std::numeric_limits<double>::epsilon() * n;
  int rank = 0;
  for (int i = 0; i < nonzero_pivots; ++i) rank += (std::fabs(lu[i * n + i]) > std::fabs(maxpivot) * threshold) ? 1 : 0;
  Vec dst(n, 0.0);
  if (rank == 0) return dst;
  // c = P rhs (row transpositions applied in order)
  Vec c(rhs);
  for (int k = 0; k < n; ++k)
    if (row_t[k] != k) std::swap(c[k], c[row_t[k]]);
  // unit lower triangular solve
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) c[i] -= lu[i * n + j] * c[j];
  // upper triangular solve on the leading rank x rank block
  for (int i = rank - 1; i >= 0; --i) {
    for (int j = i + 1; j < rank; ++j) c[i] -= lu[i * n + j] * c[j];
    c[i] /= lu[i * n + i];
  }
  // undo the column permutation: position i of the permuted unknowns is original column q[i]
  std::vector<int> q(n);
  for (int i = 0; i < n; ++i) q[i] = i;
  for (int k = 0; k < n; ++k) std::swap(q[k], q[col_t[k]]);
  for (int i = 0; i < rank; ++i) dst[q[i]] = c[i];
  return dst;
}

Vec FindInterpolatingPolynomial(const std::vector<FunctionSample>& samples) {
  const int num_samples = (int)samples.size();
  int num_constraints = 0;
  for (int i = 0; i < num_samples; ++i) {
    if (samples[i].value_is_valid) ++num_constraints;
    if (samples[i].gradient_is_valid) ++num_constraints;
  }
  const int degree = num_constraints - 1;
  std::vector<double> lhs((size_t)num_constraints * num_constraints, 0.0);
  Vec rhs(num_constraints, 0.0);
  int row = 0;
  for (int i = 0; i < num_samples; ++i) {
    const FunctionSample& sample = samples[i];
    if (sample.value_is_valid) {
      for (int j = 0; j <= degree; ++j) lhs[row * num_constraints + j] = std::pow(sample.x, degree - j);
      rhs[row] = sample.value;
      ++row;
    }
    if (sample.gradient_is_valid) {
      for (int j = 0; j < degree; ++j) lhs[row * num_constraints + j] = (degree - j) * std::pow(sample.x, degree - j - 1);
      rhs[row] = sample.gradient;
      ++row;
    }
  }
  return FullPivLuSolve(lhs, num_constraints, rhs);
}

void MinimizeInterpolatingPolynomialImpl(const std::vector<FunctionSample>& samples, double x_min, double x_max,
                                         double* optimal_x, double* optimal_value) {
  const Vec polynomial = FindInterpolatingPolynomial(samples);
  MinimizePolynomial(polynomial, x_min, x_max, optimal_x, optimal_value);
  for (size_t i = 0; i < samples.size(); ++i) {
    const FunctionSample& sample = samples[i];
    if ((sample.x < x_min) || (sample.x > x_max)) continue;
    const double value = EvaluatePolynomial(polynomial, sample.x);
    if (value < *optimal_value) {
      *optimal_x = sample.x;
      *optimal_value = value;
    }
  }
}

}  // namespace

// ================================ program / evaluator ================================
namespace {

struct ReducedProgram {
  internal::ProblemImpl* problem;
  std::vector<int> free_blocks;      // indices of the non-constant parameter blocks, program order
  std::vector<int> residual_blocks;  // indices of the residual blocks with >= 1 non-constant block
  int num_parameters;
  double fixed_cost;
  bool is_constrained;
};

bool IsFinite(double v) { return std::isfinite(v); }

// ParameterBlock::Plus for every block of the reduced program: x + delta, then projection on the box
void Plus(const ReducedProgram& P, const double* x, const double* delta, double* x_plus_delta) {
  for (size_t b = 0; b < P.free_blocks.size(); ++b) {
    const internal::ParameterBlock& pb = P.problem->parameter_blocks[P.free_blocks[b]];
    const int o = pb.reduced_offset;
    for (int i = 0; i < pb.size; ++i) x_plus_delta[o + i] = x[o + i] + delta[o + i];
    if (!pb.lower.empty())
      for (int i = 0; i < pb.size; ++i) x_plus_delta[o + i] = std::max(x_plus_delta[o + i], pb.lower[i]);
    if (!pb.upper.empty())
      for (int i = 0; i < pb.size; ++i) x_plus_delta[o + i] = std::min(x_plus_delta[o + i], pb.upper[i]);
  }
}

// Block-sparse Jacobian of the reduced program: per residual block (2 rows here, generic sizes
// kept) up to two dense blocks.
struct Jacobian {
  std::vector<double> values;            // per residual block: [block0 (nr x s0) | block1 (nr x s1)]
  std::vector<size_t> offset;            // start of each residual block's values
};

struct Evaluator {
  const ReducedProgram* P;
  std::vector<size_t> jac_offset;
  std::vector<int> res_offset;
  int num_residuals;
  size_t jac_size;
  explicit Evaluator(const ReducedProgram* p) : P(p) {
    num_residuals = 0;
    jac_size = 0;
    for (size_t k = 0; k < P->residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P->problem->residual_blocks[P->residual_blocks[k]];
      const int nr = rb.cost_function->num_residuals();
      res_offset.push_back(num_residuals);
      jac_offset.push_back(jac_size);
      num_residuals += nr;
      jac_size += (size_t)nr * (rb.cost_function->parameter_block_sizes()[0] + rb.cost_function->parameter_block_sizes()[1]);
    }
  }

  // ProgramEvaluator::Evaluate + ResidualBlock::Evaluate + Corrector.  Any of residuals,
  // gradient, jacobian may be NULL.  Cost is accumulated in residual-block order (num_threads = 1).
  bool Evaluate(const double* state, double* cost, double* residuals, double* gradient, double* jacobian) const {
    *cost = 0.0;
    const bool need_jac = gradient != NULL || jacobian != NULL;
    if (gradient) std::fill(gradient, gradient + P->num_parameters, 0.0);
    if (jacobian) std::fill(jacobian, jacobian + jac_size, 0.0);
    double scratch_r[16], scratch_j0[64], scratch_j1[64];
    for (size_t k = 0; k < P->residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P->problem->residual_blocks[P->residual_blocks[k]];
      const internal::ParameterBlock& p0 = P->problem->parameter_blocks[rb.block[0]];
      const internal::ParameterBlock& p1 = P->problem->parameter_blocks[rb.block[1]];
      const int nr = rb.cost_function->num_residuals();
      const double* params[2] = {p0.reduced_offset >= 0 ? state + p0.reduced_offset : p0.user_state,
                                 p1.reduced_offset >= 0 ? state + p1.reduced_offset : p1.user_state};
      double* r = scratch_r;
      // jacobians of constant blocks are not requested (NULL)
      double* jac[2] = {(need_jac && p0.reduced_offset >= 0) ? scratch_j0 : NULL,
                        (need_jac && p1.reduced_offset >= 0) ? scratch_j1 : NULL};
      if (!rb.cost_function->Evaluate(params, r, need_jac ? jac : NULL)) return false;
      double squared_norm = 0.0;
      for (int i = 0; i < nr; ++i) squared_norm += r[i] * r[i];
      if (rb.loss_function == NULL) {
        *cost += 0.5 * squared_norm;
      } else {
        double rho[3];
        rb.loss_function->Evaluate(squared_norm, rho);
        *cost += 0.5 * rho[0];
        if (need_jac || residuals) {
          // corrector.cc: rho'' <= 0 or zero residual -> scale by sqrt(rho'); otherwise the
          // Triggs correction.  Jacobians are corrected before the residuals.
          const double sqrt_rho1 = std::sqrt(rho[1]);
          double residual_scaling, alpha_sq_norm;
          if ((squared_norm == 0.0) || (rho[2] <= 0.0)) {
            residual_scaling = sqrt_rho1;
            alpha_sq_norm = 0.0;
          } else {
            const double D = 1.0 + 2.0 * squared_norm * rho[2] / rho[1];
            const double alpha = 1.0 - std::sqrt(D);
            residual_scaling = sqrt_rho1 / (1 - alpha);
            alpha_sq_norm = alpha / squared_norm;
          }
          for (int b = 0; b < 2; ++b) {
            if (!jac[b]) continue;
            const int sz = b == 0 ? p0.size : p1.size;
            if (alpha_sq_norm == 0.0) {
              for (int i = 0; i < nr * sz; ++i) jac[b][i] *= sqrt_rho1;
            } else {
              // J = sqrt(rho') (J - alpha / |r|^2 r r' J)
              for (int c = 0; c < sz; ++c) {
                double r_transpose_j = 0.0;
                for (int i = 0; i < nr; ++i) r_transpose_j += jac[b][i * sz + c] * r[i];
                for (int i = 0; i < nr; ++i)
                  jac[b][i * sz + c] = sqrt_rho1 * (jac[b][i * sz + c] - alpha_sq_norm * r[i] * r_transpose_j);
              }
            }
          }
          for (int i = 0; i < nr; ++i) r[i] *= residual_scaling;
        }
      }
      if (residuals)
        for (int i = 0; i < nr; ++i) residuals[res_offset[k] + i] = r[i];
      if (need_jac) {
        size_t jo = jac_offset[k];
        for (int b = 0; b < 2; ++b) {
          const internal::ParameterBlock& pb = b == 0 ? p0 : p1;
          if (jac[b]) {
            if (jacobian) std::memcpy(jacobian + jo, jac[b], sizeof(double) * nr * pb.size);
            if (gradient)  // gradient += J' r
              for (int c = 0; c < pb.size; ++c) {
                double s = 0.0;
                for (int i = 0; i < nr; ++i) s += jac[b][i * pb.size + c] * r[i];
                gradient[pb.reduced_offset + c] += s;
              }
          }
          jo += (size_t)nr * pb.size;
        }
      }
    }
    return true;
  }
};

// ================================ linear solver ================================
// SparseNormalCholeskySolver: (J'J + D'D) y = J'r, factorised exactly.  Here: an envelope
// (skyline) Cholesky in the given variable order — the fill-reducing order Ceres asks
// SuiteSparse for only changes round-off.  Returns false when a pivot is not positive.
struct NormalEquations {
  int n;
  std::vector<int> first;       // first[i] = column of the first stored entry of row i (envelope)
  std::vector<size_t> rowptr;   // row i is stored at vals[rowptr[i] .. rowptr[i] + (i - first[i]) ]
  std::vector<double> vals;
  double& at(int i, int j) { return vals[rowptr[i] + (size_t)(j - first[i])]; }  // j <= i, j >= first[i]
};

}  // namespace

// ================================ trust_region_minimizer.cc ================================
namespace {

std::mutex g_record_mutex;
bool g_recording = false;
std::vector<shim::Record> g_records;

struct Minimizer {
  const Solver::Options& options;
  const ReducedProgram& P;
  Evaluator evaluator;
  Solver::Summary* summary;
  int n;
  // state
  Vec x, candidate_x, gradient, delta, trust_region_step, jacobian_scaling, residuals, model_residuals;
  Vec negative_gradient, jacobian, diagonal, lm_diagonal;
  double x_norm, x_cost, minimum_cost, candidate_cost, model_cost_change;
  double gradient_max_norm;
  int iteration;
  bool step_is_valid, step_is_successful;
  int num_consecutive_invalid_steps;
  double* parameters;
  // LevenbergMarquardtStrategy
  double radius, decrease_factor;
  bool reuse_diagonal;
  // TrustRegionStepEvaluator (max_consecutive_nonmonotonic_steps = 0)
  double se_minimum_cost, se_current_cost, se_reference_cost, se_candidate_cost;
  double se_acc_reference_mcc, se_acc_candidate_mcc;
  int se_num_nonmonotonic;
  // envelope of J'J (structure is fixed)
  NormalEquations ne;

  Minimizer(const Solver::Options& o, const ReducedProgram& p, Solver::Summary* s)
      : options(o), P(p), evaluator(&p), summary(s), n(p.num_parameters) {}

  void BuildEnvelope() {
    ne.n = n;
    ne.first.resize(n);
    for (int i = 0; i < n; ++i) ne.first[i] = i;
    for (size_t k = 0; k < P.residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P.problem->residual_blocks[P.residual_blocks[k]];
      const internal::ParameterBlock& p0 = P.problem->parameter_blocks[rb.block[0]];
      const internal::ParameterBlock& p1 = P.problem->parameter_blocks[rb.block[1]];
      int lo = n;
      if (p0.reduced_offset >= 0) lo = std::min(lo, p0.reduced_offset);
      if (p1.reduced_offset >= 0) lo = std::min(lo, p1.reduced_offset);
      const internal::ParameterBlock* pbs[2] = {&p0, &p1};
      for (int b = 0; b < 2; ++b) {
        if (pbs[b]->reduced_offset < 0) continue;
        for (int i = 0; i < pbs[b]->size; ++i) ne.first[pbs[b]->reduced_offset + i] = std::min(ne.first[pbs[b]->reduced_offset + i], lo);
      }
    }
    ne.rowptr.resize(n + 1);
    size_t tot = 0;
    for (int i = 0; i < n; ++i) {
      ne.rowptr[i] = tot;
      tot += (size_t)(i - ne.first[i] + 1);
    }
    ne.rowptr[n] = tot;
    ne.vals.assign(tot, 0.0);
  }

  // lhs = J'J + D'D (lower envelope), rhs = J'r; Cholesky; y.  Returns false on failure.
  bool SolveNormalEquations(const double* D, double* y) {
    std::fill(ne.vals.begin(), ne.vals.end(), 0.0);
    Vec rhs(n, 0.0);
    for (size_t k = 0; k < P.residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P.problem->residual_blocks[P.residual_blocks[k]];
      const internal::ParameterBlock* pbs[2] = {&P.problem->parameter_blocks[rb.block[0]],
                                                &P.problem->parameter_blocks[rb.block[1]]};
      const int nr = rb.cost_function->num_residuals();
      const double* r = &residuals[evaluator.res_offset[k]];
      const double* jb[2];
      jb[0] = &jacobian[evaluator.jac_offset[k]];
      jb[1] = jb[0] + (size_t)nr * pbs[0]->size;
      for (int a = 0; a < 2; ++a) {
        if (pbs[a]->reduced_offset < 0) continue;
        const int sa = pbs[a]->size, oa = pbs[a]->reduced_offset;
        for (int ca = 0; ca < sa; ++ca) {
          double s = 0.0;
          for (int i = 0; i < nr; ++i) s += jb[a][i * sa + ca] * r[i];
          rhs[oa + ca] += s;
        }
        for (int b = 0; b < 2; ++b) {
          if (pbs[b]->reduced_offset < 0) continue;
          const int sb = pbs[b]->size, ob = pbs[b]->reduced_offset;
          for (int ca = 0; ca < sa; ++ca)
            for (int cb = 0; cb < sb; ++cb) {
              const int row = oa + ca, col = ob + cb;
              if (col > row) continue;  // lower triangle
              double s = 0.0;
              for (int i = 0; i < nr; ++i) s += jb[a][i * sa + ca] * jb[b][i * sb + cb];
              ne.at(row, col) += s;
            }
        }
      }
    }
    for (int i = 0; i < n; ++i) ne.at(i, i) += D[i] * D[i];
    // envelope Cholesky: L L' = A, row by row
    for (int i = 0; i < n; ++i) {
      const int fi = ne.first[i];
      for (int j = fi; j <= i; ++j) {
        const int fj = ne.first[j];
        double s = ne.at(i, j);
        for (int k = std::max(fi, fj); k < j; ++k) s -= ne.at(i, k) * ne.at(j, k);
        if (j < i) {
          ne.at(i, j) = s / ne.at(j, j);
        } else {
          if (!(s > 0.0) || !IsFinite(s)) return false;
          ne.at(i, i) = std::sqrt(s);
        }
      }
    }
    for (int i = 0; i < n; ++i) {
      double s = rhs[i];
      for (int k = ne.first[i]; k < i; ++k) s -= ne.at(i, k) * y[k];
      y[i] = s / ne.at(i, i);
    }
    for (int i = n - 1; i >= 0; --i) {
      y[i] /= ne.at(i, i);
      const double yi = y[i];
      for (int k = ne.first[i]; k < i; ++k) y[k] -= ne.at(i, k) * yi;
    }
    return true;
  }

  static double Norm(const Vec& v) {
    double s = 0.0;
    for (size_t i = 0; i < v.size(); ++i) s += v[i] * v[i];
    return std::sqrt(s);
  }

  void SquaredColumnNorm(double* out) const {
    std::fill(out, out + n, 0.0);
    for (size_t k = 0; k < P.residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P.problem->residual_blocks[P.residual_blocks[k]];
      const int nr = rb.cost_function->num_residuals();
      size_t jo = evaluator.jac_offset[k];
      for (int b = 0; b < 2; ++b) {
        const internal::ParameterBlock& pb = P.problem->parameter_blocks[rb.block[b]];
        if (pb.reduced_offset >= 0)
          for (int i = 0; i < nr; ++i)
            for (int c = 0; c < pb.size; ++c) {
              const double v = jacobian[jo + (size_t)i * pb.size + c];
              out[pb.reduced_offset + c] += v * v;
            }
        jo += (size_t)nr * pb.size;
      }
    }
  }

  void ScaleColumns(const double* scale) {
    for (size_t k = 0; k < P.residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P.problem->residual_blocks[P.residual_blocks[k]];
      const int nr = rb.cost_function->num_residuals();
      size_t jo = evaluator.jac_offset[k];
      for (int b = 0; b < 2; ++b) {
        const internal::ParameterBlock& pb = P.problem->parameter_blocks[rb.block[b]];
        if (pb.reduced_offset >= 0)
          for (int i = 0; i < nr; ++i)
            for (int c = 0; c < pb.size; ++c) jacobian[jo + (size_t)i * pb.size + c] *= scale[pb.reduced_offset + c];
        jo += (size_t)nr * pb.size;
      }
    }
  }

  // model_residuals = J step
  void RightMultiply(const double* v, double* out) const {
    std::fill(out, out + evaluator.num_residuals, 0.0);
    for (size_t k = 0; k < P.residual_blocks.size(); ++k) {
      const internal::ResidualBlock& rb = P.problem->residual_blocks[P.residual_blocks[k]];
      const int nr = rb.cost_function->num_residuals();
      size_t jo = evaluator.jac_offset[k];
      for (int b = 0; b < 2; ++b) {
        const internal::ParameterBlock& pb = P.problem->parameter_blocks[rb.block[b]];
        if (pb.reduced_offset >= 0)
          for (int i = 0; i < nr; ++i)
            for (int c = 0; c < pb.size; ++c)
              out[evaluator.res_offset[k] + i] += jacobian[jo + (size_t)i * pb.size + c] * v[pb.reduced_offset + c];
        jo += (size_t)nr * pb.size;
      }
    }
  }

  bool EvaluateGradientAndJacobian() {
    if (!evaluator.Evaluate(x.data(), &x_cost, residuals.data(), gradient.data(), jacobian.data())) {
      summary->message = "Residual and Jacobian evaluation failed.";
      return false;
    }
    if (P.is_constrained) {
      // the projected gradient:  x - Plus(x, -gradient)
      for (int i = 0; i < n; ++i) delta[i] = -gradient[i];
      Plus(P, x.data(), delta.data(), negative_gradient.data());
      gradient_max_norm = 0.0;
      for (int i = 0; i < n; ++i) gradient_max_norm = std::max(gradient_max_norm, std::fabs(x[i] - negative_gradient[i]));
    } else {
      gradient_max_norm = 0.0;
      for (int i = 0; i < n; ++i) gradient_max_norm = std::max(gradient_max_norm, std::fabs(gradient[i]));
    }
    if (options.jacobi_scaling) {
      if (iteration == 0) {
        // computed once, with one added to the denominator to prevent division by zero
        SquaredColumnNorm(jacobian_scaling.data());
        for (int i = 0; i < n; ++i) jacobian_scaling[i] = 1.0 / (1.0 + std::sqrt(jacobian_scaling[i]));
      }
      ScaleColumns(jacobian_scaling.data());
    }
    return true;
  }

  // line_search.cc: LineSearchFunction::Evaluate
  void LineSearchEvaluate(const Vec& position, const Vec& direction, double step, FunctionSample* out, Vec* scratch_x,
                          Vec* scratch_g) const {
    *out = FunctionSample();
    out->x = step;
    Vec scaled(n);
    for (int i = 0; i < n; ++i) scaled[i] = out->x * direction[i];
    Plus(P, position.data(), scaled.data(), scratch_x->data());
    const bool ok = evaluator.Evaluate(scratch_x->data(), &out->value, NULL, scratch_g->data(), NULL);
    if (!ok || !IsFinite(out->value)) return;
    out->value_is_valid = true;
    double g = 0.0;
    for (int i = 0; i < n; ++i) g += direction[i] * (*scratch_g)[i];
    out->gradient = g;
    if (!IsFinite(out->gradient)) return;
    out->gradient_is_valid = true;
  }

  // line_search.cc: LineSearch::InterpolatingPolynomialMinimizingStepSize (CUBIC)
  static double InterpolatingPolynomialMinimizingStepSize(const FunctionSample& lowerbound, const FunctionSample& previous,
                                                          const FunctionSample& current, double min_step_size,
                                                          double max_step_size) {
    if (!current.value_is_valid) return std::min(std::max(current.x * 0.5, min_step_size), max_step_size);
    std::vector<FunctionSample> samples;
    samples.push_back(lowerbound);
    samples.push_back(current);  // two point interpolation using the function values and the gradients
    if (previous.value_is_valid) samples.push_back(previous);  // three point interpolation
    double step_size = 0.0, unused_min_value = 0.0;
    MinimizeInterpolatingPolynomialImpl(samples, min_step_size, max_step_size, &step_size, &unused_min_value);
    return step_size;
  }

  // trust_region_minimizer.cc: DoLineSearch + line_search.cc: ArmijoLineSearch::DoSearch
  void DoLineSearch() {
    FunctionSample initial_position;
    initial_position.x = 0.0;
    initial_position.value = x_cost;
    initial_position.value_is_valid = true;
    double initial_gradient = 0.0;
    for (int i = 0; i < n; ++i) initial_gradient += gradient[i] * delta[i];
    initial_position.gradient = initial_gradient;
    initial_position.gradient_is_valid = true;
    double descent_direction_max_norm = 0.0;
    for (int i = 0; i < n; ++i) descent_direction_max_norm = std::max(descent_direction_max_norm, std::fabs(delta[i]));
    FunctionSample previous, current;
    Vec sx(n), sg(n);
    const Vec direction(delta);
    int num_iterations = 0;
    LineSearchEvaluate(x, direction, 1.0, &current, &sx, &sg);
    bool success = true;
    while (!current.value_is_valid ||
           current.value > (x_cost + options.line_search_sufficient_function_decrease * initial_gradient * current.x)) {
      ++num_iterations;
      if (num_iterations >= options.max_num_line_search_step_size_iterations) {
        success = false;  // "Armijo failed to find a point satisfying the sufficient decrease condition"
        break;
      }
      const double step_size = InterpolatingPolynomialMinimizingStepSize(
          initial_position, previous, current, options.max_line_search_step_contraction * current.x,
          options.min_line_search_step_contraction * current.x);
      if (step_size * descent_direction_max_norm < options.min_line_search_step_size) {
        success = false;  // "step_size too small"
        break;
      }
      previous = current;
      LineSearchEvaluate(x, direction, step_size, &current, &sx, &sg);
    }
    summary->num_line_search_steps += num_iterations;
    if (success)
      for (int i = 0; i < n; ++i) delta[i] *= current.x;
  }

  // TrustRegionStepEvaluator
  double StepQuality(double cost, double mcc) const {
    const double relative_decrease = (se_current_cost - cost) / mcc;
    const double historical_relative_decrease = (se_reference_cost - cost) / (se_acc_reference_mcc + mcc);
    return std::max(relative_decrease, historical_relative_decrease);
  }
  void StepAcceptedEvaluator(double cost, double mcc) {
    se_current_cost = cost;
    se_acc_candidate_mcc += mcc;
    se_acc_reference_mcc += mcc;
    if (se_current_cost < se_minimum_cost) {
      se_minimum_cost = se_current_cost;
      se_num_nonmonotonic = 0;
      se_candidate_cost = se_current_cost;
      se_acc_candidate_mcc = 0.0;
    } else {
      ++se_num_nonmonotonic;
      if (se_current_cost > se_candidate_cost) {
        se_candidate_cost = se_current_cost;
        se_acc_candidate_mcc = 0.0;
      }
    }
    if (se_num_nonmonotonic == 0 /* max_consecutive_nonmonotonic_steps */) {
      se_reference_cost = se_candidate_cost;
      se_acc_reference_mcc = se_acc_candidate_mcc;
    }
  }

  void Minimize(double* params) {
    parameters = params;
    x.assign(params, params + n);
    candidate_x.assign(n, 0.0);
    gradient.assign(n, 0.0);
    delta.assign(n, 0.0);
    trust_region_step.assign(n, 0.0);
    jacobian_scaling.assign(n, 1.0);
    negative_gradient.assign(n, 0.0);
    residuals.assign(evaluator.num_residuals, 0.0);
    model_residuals.assign(evaluator.num_residuals, 0.0);
    jacobian.assign(evaluator.jac_size, 0.0);
    diagonal.assign(n, 0.0);
    lm_diagonal.assign(n, 0.0);
    BuildEnvelope();
    x_norm = Norm(x);
    minimum_cost = std::numeric_limits<double>::max();
    num_consecutive_invalid_steps = 0;
    radius = options.initial_trust_region_radius;
    decrease_factor = 2.0;
    reuse_diagonal = false;
    summary->num_successful_steps = 0;
    summary->num_line_search_steps = 0;

    // ---- IterationZero
    iteration = 0;
    if (P.is_constrained) {
      std::fill(delta.begin(), delta.end(), 0.0);
      Plus(P, x.data(), delta.data(), candidate_x.data());
      x = candidate_x;
      x_norm = Norm(x);
    }
    if (!EvaluateGradientAndJacobian()) {
      summary->termination_type = FAILURE;
      return;
    }
    summary->initial_cost = x_cost + summary->fixed_cost;
    step_is_valid = true;
    step_is_successful = true;
    se_minimum_cost = se_current_cost = se_reference_cost = se_candidate_cost = x_cost;
    se_acc_reference_mcc = se_acc_candidate_mcc = 0.0;
    se_num_nonmonotonic = 0;

    for (;;) {
      // ---- FinalizeIterationAndCheckIfMinimizerCanContinue
      if (step_is_successful) {
        ++summary->num_successful_steps;
        if (x_cost < minimum_cost) {
          minimum_cost = x_cost;
          std::copy(x.begin(), x.end(), parameters);
        }
      }
      summary->num_iterations = iteration;
      if (iteration >= options.max_num_iterations) {
        summary->message = "Maximum number of iterations reached.";
        summary->termination_type = NO_CONVERGENCE;
        return;
      }
      if (step_is_successful && gradient_max_norm <= options.gradient_tolerance) {
        summary->message = "Gradient tolerance reached.";
        summary->termination_type = CONVERGENCE;
        return;
      }
      if (radius <= options.min_trust_region_radius) {
        summary->message = "Minimum trust region radius reached.";
        summary->termination_type = CONVERGENCE;
        return;
      }
      ++iteration;
      step_is_valid = false;
      step_is_successful = false;

      // ---- ComputeTrustRegionStep (LevenbergMarquardtStrategy::ComputeStep)
      if (!reuse_diagonal) {
        SquaredColumnNorm(diagonal.data());
        for (int i = 0; i < n; ++i) diagonal[i] = std::min(std::max(diagonal[i], options.min_lm_diagonal), options.max_lm_diagonal);
      }
      for (int i = 0; i < n; ++i) lm_diagonal[i] = std::sqrt(diagonal[i] / radius);
      bool solved = SolveNormalEquations(lm_diagonal.data(), trust_region_step.data());
      if (solved)
        for (int i = 0; i < n; ++i)
          if (!IsFinite(trust_region_step[i])) solved = false;
      if (solved)
        for (int i = 0; i < n; ++i) trust_region_step[i] *= -1.0;
      reuse_diagonal = true;
      if (solved) {
        // model_cost_change = -(J step)' (f + J step / 2)
        RightMultiply(trust_region_step.data(), model_residuals.data());
        double dot = 0.0;
        for (int i = 0; i < evaluator.num_residuals; ++i) dot += model_residuals[i] * (residuals[i] + model_residuals[i] / 2.0);
        model_cost_change = -dot;
        step_is_valid = (model_cost_change > 0.0);
        if (step_is_valid) {
          for (int i = 0; i < n; ++i) delta[i] = trust_region_step[i] * jacobian_scaling[i];  // undo the column scaling
          num_consecutive_invalid_steps = 0;
        }
      }
      if (!step_is_valid) {
        // ---- HandleInvalidStep
        ++num_consecutive_invalid_steps;
        if (num_consecutive_invalid_steps >= options.max_num_consecutive_invalid_steps) {
          summary->message = "Number of consecutive invalid steps more than Solver::Options::max_num_consecutive_invalid_steps";
          summary->termination_type = FAILURE;
          return;
        }
        radius = radius / decrease_factor;  // StepIsInvalid = StepRejected
        decrease_factor *= 2.0;
        reuse_diagonal = true;
        continue;
      }
      if (P.is_constrained) DoLineSearch();  // projected line search enforcing the bounds

      // ---- ComputeCandidatePointAndEvaluateCost
      Plus(P, x.data(), delta.data(), candidate_x.data());
      if (!evaluator.Evaluate(candidate_x.data(), &candidate_cost, NULL, NULL, NULL) || !IsFinite(candidate_cost))
        candidate_cost = std::numeric_limits<double>::max();

      // ---- ParameterToleranceReached
      double step_norm = 0.0;
      for (int i = 0; i < n; ++i) step_norm += (x[i] - candidate_x[i]) * (x[i] - candidate_x[i]);
      step_norm = std::sqrt(step_norm);
      if (step_norm <= options.parameter_tolerance * (x_norm + options.parameter_tolerance)) {
        summary->message = "Parameter tolerance reached.";
        summary->termination_type = CONVERGENCE;
        summary->num_iterations = iteration;
        return;
      }
      // ---- FunctionToleranceReached
      const double cost_change = x_cost - candidate_cost;
      if (std::fabs(cost_change) <= options.function_tolerance * x_cost) {
        summary->message = "Function tolerance reached.";
        summary->termination_type = CONVERGENCE;
        summary->num_iterations = iteration;
        return;
      }
      // ---- IsStepSuccessful
      const double relative_decrease = StepQuality(candidate_cost, model_cost_change);
      if (relative_decrease > options.min_relative_decrease) {
        // ---- HandleSuccessfulStep
        x = candidate_x;
        x_norm = Norm(x);
        if (!EvaluateGradientAndJacobian()) {
          summary->termination_type = FAILURE;
          return;
        }
        step_is_successful = true;
        radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * relative_decrease - 1.0, 3));
        radius = std::min(options.max_trust_region_radius, radius);
        decrease_factor = 2.0;
        reuse_diagonal = false;
        StepAcceptedEvaluator(candidate_cost, model_cost_change);
      } else {
        // ---- HandleUnsuccessfulStep
        radius = radius / decrease_factor;
        decrease_factor *= 2.0;
        reuse_diagonal = true;
      }
    }
  }
};

}  // namespace

// ================================ solver.cc ================================
void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary) {
  internal::ProblemImpl* impl = problem->impl();
  *summary = Solver::Summary();
  // ---- preprocessing: remove constant parameter blocks and the residual blocks that depend
  // only on them (their cost becomes fixed_cost) — Program::RemoveFixedBlocks
  ReducedProgram P;
  P.problem = impl;
  P.num_parameters = 0;
  P.fixed_cost = 0.0;
  P.is_constrained = false;
  std::vector<char> used(impl->parameter_blocks.size(), 0);
  for (size_t k = 0; k < impl->residual_blocks.size(); ++k) {
    const internal::ResidualBlock& rb = impl->residual_blocks[k];
    const bool c0 = impl->parameter_blocks[rb.block[0]].is_constant, c1 = impl->parameter_blocks[rb.block[1]].is_constant;
    if (c0 && c1) {
      // fixed cost: evaluated once at the user state
      const double* params[2] = {impl->parameter_blocks[rb.block[0]].user_state, impl->parameter_blocks[rb.block[1]].user_state};
      double r[16];
      if (rb.cost_function->Evaluate(params, r, NULL)) {
        double sq = 0.0;
        for (int i = 0; i < rb.cost_function->num_residuals(); ++i) sq += r[i] * r[i];
        if (rb.loss_function) {
          double rho[3];
          rb.loss_function->Evaluate(sq, rho);
          P.fixed_cost += 0.5 * rho[0];
        } else {
          P.fixed_cost += 0.5 * sq;
        }
      }
      continue;
    }
    P.residual_blocks.push_back((int)k);
    used[rb.block[0]] = 1;
    used[rb.block[1]] = 1;
  }
  for (size_t b = 0; b < impl->parameter_blocks.size(); ++b) {
    internal::ParameterBlock& pb = impl->parameter_blocks[b];
    pb.reduced_offset = -1;
    if (pb.is_constant || !used[b]) continue;
    pb.reduced_offset = P.num_parameters;
    P.num_parameters += pb.size;
    P.free_blocks.push_back((int)b);
    if (!pb.lower.empty() || !pb.upper.empty()) P.is_constrained = true;
  }
  summary->fixed_cost = P.fixed_cost;
  summary->num_parameter_blocks_reduced = (int)P.free_blocks.size();
  summary->num_residual_blocks_reduced = (int)P.residual_blocks.size();
  if (P.free_blocks.empty()) {
    // "Function tolerance reached. No non-constant parameter blocks found."
    summary->message = "Function tolerance reached. No non-constant parameter blocks found.";
    summary->termination_type = CONVERGENCE;
    summary->initial_cost = summary->final_cost = P.fixed_cost;
  } else {
    Vec reduced(P.num_parameters);
    for (size_t b = 0; b < P.free_blocks.size(); ++b) {
      const internal::ParameterBlock& pb = impl->parameter_blocks[P.free_blocks[b]];
      for (int i = 0; i < pb.size; ++i) reduced[pb.reduced_offset + i] = pb.user_state[i];
    }
    const Vec original(reduced);
    Minimizer minimizer(options, P, summary);
    minimizer.Minimize(reduced.data());
    // solver.cc: the user state is only updated when the solution is usable
    const Vec& result = summary->IsSolutionUsable() ? reduced : original;
    for (size_t b = 0; b < P.free_blocks.size(); ++b) {
      const internal::ParameterBlock& pb = impl->parameter_blocks[P.free_blocks[b]];
      for (int i = 0; i < pb.size; ++i) pb.user_state[i] = result[pb.reduced_offset + i];
    }
    if (summary->IsSolutionUsable()) summary->final_cost = minimizer.minimum_cost + P.fixed_cost;
  }
  if (g_recording) {
    std::lock_guard<std::mutex> lock(g_record_mutex);
    shim::Record rec;
    rec.first_parameter_block = impl->parameter_blocks.empty() ? NULL : impl->parameter_blocks[0].user_state;
    for (size_t b = 0; b < impl->parameter_blocks.size(); ++b)
      if (impl->parameter_blocks[b].user_state < rec.first_parameter_block) rec.first_parameter_block = impl->parameter_blocks[b].user_state;
    rec.summary = *summary;
    g_records.push_back(rec);
  }
}

namespace shim {

void SetTukeyVariant(int variant) { g_tukey_variant = variant == 2 ? 2 : 1; }
void SetRecording(bool on) {
  std::lock_guard<std::mutex> lock(g_record_mutex);
  g_recording = on;
}
std::vector<Record> TakeRecords() {
  std::lock_guard<std::mutex> lock(g_record_mutex);
  std::vector<Record> out;
  out.swap(g_records);
  return out;
}

void MinimizeInterpolatingPolynomial(const double* s5, int n_samples, double x_min, double x_max, double* optimal_x,
                                     double* optimal_value) {
  std::vector<FunctionSample> samples(n_samples);
  for (int i = 0; i < n_samples; ++i) {
    samples[i].x = s5[5 * i];
    samples[i].value = s5[5 * i + 1];
    samples[i].gradient = s5[5 * i + 2];
    samples[i].value_is_valid = s5[5 * i + 3] != 0.0;
    samples[i].gradient_is_valid = s5[5 * i + 4] != 0.0;
  }
  MinimizeInterpolatingPolynomialImpl(samples, x_min, x_max, optimal_x, optimal_value);
}

int FindPolynomialRoots(const double* poly, int n_coeff, double* real, double* imag) {
  Vec r, im;
  if (!FindPolynomialRootsImpl(Vec(poly, poly + n_coeff), &r, &im)) return -1;
  for (size_t i = 0; i < r.size(); ++i) {
    real[i] = r[i];
    if (imag) imag[i] = im[i];
  }
  return (int)r.size();
}

}  // namespace shim

// Recording can also be switched on from the environment (the reference's main() cannot be
// edited): LFR_CERES_RECORD_FILE=<path> makes the process dump one line per ceres::Solve at exit.
namespace {
struct RecordDumper {
  RecordDumper() {
    if (std::getenv("LFR_CERES_RECORD_FILE")) g_recording = true;
  }
  ~RecordDumper() {
    const char* path = std::getenv("LFR_CERES_RECORD_FILE");
    if (!path) return;
    std::FILE* f = std::fopen(path, "w");
    if (!f) return;
    for (size_t i = 0; i < g_records.size(); ++i) {
      const shim::Record& r = g_records[i];
      std::fprintf(f, "%llu %d %d %d %.17g %.17g %d %d %.17g\n", (unsigned long long)(size_t)r.first_parameter_block,
                   r.summary.num_iterations, (int)r.summary.termination_type, r.summary.num_line_search_steps,
                   r.summary.initial_cost, r.summary.final_cost, r.summary.num_parameter_blocks_reduced,
                   r.summary.num_residual_blocks_reduced, r.summary.fixed_cost);
    }
    std::fclose(f);
  }
};
RecordDumper g_record_dumper;
}  // namespace

}  // namespace ceres
