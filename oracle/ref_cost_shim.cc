// ref_cost_shim.cc — C entry points into the REFERENCE's own cost.cc.
//
// TEST INFRASTRUCTURE ONLY.  This translation unit #includes
// /root/reference/multi-view-refinement/cost.cc verbatim (found through -I; never copied into
// this repository) and compiles it against the shim headers in oracle/ref_shims/ (Eigen/Core,
// ceres/ceres.h).  What runs below is therefore the reference author's
// BiquadraticInterpolator::Evaluate (cost.cc:13-48), its Jet overload (cost.cc:56-63) and
// InterpolatedCostFunctor (cost.cc:78-94), differentiated by a Jet type that follows
// ceres/jet.h's published definition.  Built by oracle/build_ref.py into
// oracle/_ref/libref_cost.so; used by tests/test_ref_cost.py to pin the restated oracle
// (lfr_ref_interpolate, lfr_debug_edge_eval in oracle/lfr_oracle.cc) and the CUDA kernel.
#include "cost.cc"

#include <stdint.h>

extern "C" {

// cost.cc:13-48 on n inputs: grids [n][18] doubles, rc [n][2]; out f [n][2], dfdrow [n][2], dfdcol [n][2]
void lfr_refsrc_interpolate(uint64_t n, const double* grids, const double* rc, double* f, double* dfdrow,
                            double* dfdcol) {
  for (uint64_t e = 0; e < n; ++e) {
    const std::vector<double> data(grids + 18 * e, grids + 18 * e + 18);
    const BiquadraticInterpolator interpolator(data, 2);  // solve.cc:103
    interpolator.Evaluate(rc[2 * e], rc[2 * e + 1], f + 2 * e, dfdrow + 2 * e, dfdcol + 2 * e);
  }
}

// The residual block of solve.cc:107-113 without its loss: InterpolatedCostFunctor<...>::Create
// (cost.cc:92-94) evaluated through ceres::CostFunction::Evaluate.  x1, x2 [n][2]; out
// residuals [n][2], jac1 / jac2 [n][4] row-major 2x2 (d r / d x1, d r / d x2).
void lfr_refsrc_cost(uint64_t n, const double* grids, const double* x1, const double* x2, double* residuals,
                     double* jac1, double* jac2) {
  for (uint64_t e = 0; e < n; ++e) {
    const std::vector<double> data(grids + 18 * e, grids + 18 * e + 18);
    BiquadraticInterpolator interpolator(data, 2);
    ceres::CostFunction* cost = InterpolatedCostFunctor<BiquadraticInterpolator>::Create(std::move(interpolator));
    const double* params[2] = {x1 + 2 * e, x2 + 2 * e};
    double* jacs[2] = {jac1 + 4 * e, jac2 + 4 * e};
    cost->Evaluate(params, residuals + 2 * e, jacs);
    delete cost;
  }
}

// residual only (the double instantiation of operator(), cost.cc:51-53)
void lfr_refsrc_residual(uint64_t n, const double* grids, const double* x1, const double* x2, double* residuals) {
  for (uint64_t e = 0; e < n; ++e) {
    const std::vector<double> data(grids + 18 * e, grids + 18 * e + 18);
    BiquadraticInterpolator interpolator(data, 2);
    ceres::CostFunction* cost = InterpolatedCostFunctor<BiquadraticInterpolator>::Create(std::move(interpolator));
    const double* params[2] = {x1 + 2 * e, x2 + 2 * e};
    cost->Evaluate(params, residuals + 2 * e, NULL);
    delete cost;
  }
}

}  // extern "C"
