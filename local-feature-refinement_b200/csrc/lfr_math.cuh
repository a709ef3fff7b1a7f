// lfr_math.cuh — per-edge arithmetic of the multi-view refinement cost, fp64.
//
//   eval_edge()         cost.cc:13-48 (biquadratic interpolation with clamping)
//                       + cost.cc:78-90 (residual x2 - x1 - flow(x1))
//                       + Ceres loss (ScaledLoss(Cauchy|Tukey))
//   ls_next_step()      Ceres ArmijoLineSearch step-size selection by cubic /
//                       quintic Hermite interpolation (polynomial.cc)
//
// The flow grid is read as five 128-bit words (the 80-byte lfr_edge record) and
// widened to fp64 in registers; no tensor cores: this is 2-DoF sparse NLLS.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/lfr.h"

namespace lfr {

struct DevConsts {
  double bound;
  double cauchy_b, cauchy_c;          // a^2, 1/a^2
  double tukey_a2, tukey_inv_a2;      // a^2, 1/a^2
  double tukey_rho0, tukey_rho1;      // a^2/6 & 0.5 (Ceres 1.x)  or  a^2/3 & 1.0 (2.x)
  double f_tol, g_tol, p_tol;
  double radius0, radius_max, radius_min;
  double min_rel_decrease, min_diag, max_diag;
  double ls_suff, ls_max_contraction, ls_min_contraction, ls_min_step;
  int max_iter, max_invalid, max_ls_iter, linear_solver;
};

// Result of one residual-block evaluation, already in the form the assembly
// needs:  a = sim * rho'(|r|^2)  (Corrector: J~ = sqrt(a) J, r~ = sqrt(a) r),
// r = raw residual, M = I + grad(flow)  (so d r/d x_src = -M, d r/d x_dst = I).
struct EdgeEval {
  double a, r0, r1, m00, m01, m10, m11, half_rho;
};

__device__ __forceinline__ void lagrange3(double t, double L[3], double dL[3]) {
  // cost.cc:20-23, nodes at -0.5, 0, +0.5
  L[0] = 2. * t * (t - .5);
  L[1] = (-4.) * (t - .5) * (t + .5);
  L[2] = 2. * t * (t + .5);
  dL[0] = 4. * t - 1.;
  dL[1] = -8. * t;
  dL[2] = 4. * t + 1.;
}

// Scaled loss {0.5*rho, rho'} at s = |r|^2.
__device__ __forceinline__ void scaled_loss(int kind, double sim, double s, const DevConsts& K,
                                            double* half_rho, double* rho1) {
  if (kind == LFR_EDGE_CAUCHY) {
    const double sum = 1.0 + s * K.cauchy_c;
    *half_rho = 0.5 * sim * (K.cauchy_b * log(sum));
    *rho1 = sim * fmax(2.2250738585072014e-308, 1.0 / sum);
  } else {
    if (s <= K.tukey_a2) {
      const double v = 1.0 - s * K.tukey_inv_a2;
      const double v2 = v * v;
      *half_rho = 0.5 * sim * (K.tukey_rho0 * (1.0 - v2 * v));
      *rho1 = sim * (K.tukey_rho1 * v2);
    } else {
      *half_rho = 0.5 * sim * K.tukey_rho0;
      *rho1 = 0.0;
    }
  }
}

// q[0..4] = the 80-byte edge record as five float4 (q[4].z = sim, q[4].w = dst).
__device__ __forceinline__ EdgeEval eval_edge(const float4 q[5], int kind, double x1r, double x1c,
                                              double x2r, double x2c, const DevConsts& K) {
  const double row = fmax(fmin(x1r, .5), -.5);   // cost.cc:17-18
  const double col = fmax(fmin(x1c, .5), -.5);
  const bool row_free = (row == x1r), col_free = (col == x1c);  // cost.cc:38,41
  double Lr[3], dLr[3], Lc[3], dLc[3];
  lagrange3(row, Lr, dLr);
  lagrange3(col, Lc, dLc);
  // D[i][j][k] = flow[2*(3i+j)+k]
  const double D[3][3][2] = {
      {{(double)q[0].x, (double)q[0].y}, {(double)q[0].z, (double)q[0].w}, {(double)q[1].x, (double)q[1].y}},
      {{(double)q[1].z, (double)q[1].w}, {(double)q[2].x, (double)q[2].y}, {(double)q[2].z, (double)q[2].w}},
      {{(double)q[3].x, (double)q[3].y}, {(double)q[3].z, (double)q[3].w}, {(double)q[4].x, (double)q[4].y}}};
  double f[2], fr[2], fc[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    double v = 0., vr = 0., vc = 0.;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double t = Lc[0] * D[i][0][k] + Lc[1] * D[i][1][k] + Lc[2] * D[i][2][k];
      const double tc = dLc[0] * D[i][0][k] + dLc[1] * D[i][1][k] + dLc[2] * D[i][2][k];
      v += Lr[i] * t;
      vr += dLr[i] * t;
      vc += Lr[i] * tc;
    }
    f[k] = v;
    fr[k] = row_free ? vr : 0.;
    fc[k] = col_free ? vc : 0.;
  }
  EdgeEval o;
  o.r0 = x2r - x1r - f[0];   // cost.cc:87
  o.r1 = x2c - x1c - f[1];
  o.m00 = 1. + fr[0];
  o.m01 = fc[0];
  o.m10 = fr[1];
  o.m11 = 1. + fc[1];
  scaled_loss(kind, (double)q[4].z, o.r0 * o.r0 + o.r1 * o.r1, K, &o.half_rho, &o.a);
  return o;
}

// ---------------------------------------------------------------------------
// Line-search step selection (Ceres ArmijoLineSearch, polynomial.cc).
//
// Ceres fits the polynomial interpolating {value, gradient} at step 0, at the
// current trial step and (from the second contraction on) at the previous one,
// and takes the minimiser over [1e-3, 0.6] x current among: interval middle,
// interval ends, real parts of all roots of p', and the sample abscissae.  The
// same interpolant is built here in the normalised variable t = x / h
// (h = largest sample step), where the two constraints at 0 fix the two lowest
// coefficients and the rest is a 2x2 (cubic) or 4x4 (quintic) solve — the same
// polynomial as Ceres' 4x4 / 6x6 Vandermonde solve, better conditioned.  The
// quartic p' is solved by an Aberth-Ehrlich iteration (= eigenvalues of Ceres'
// companion matrix) with one root per lane.  oracle/lfr_oracle.cc mirrors this
// operation for operation.
// ---------------------------------------------------------------------------
struct LsSample {
  double x, value, gradient;
  bool value_valid, gradient_valid;
};

__device__ __forceinline__ double poly_eval(const double* p, int n, double x) {
  double v = 0.0;
  for (int i = 0; i < n; ++i) v = v * x + p[i];
  return v;
}

// Real parts of the roots of the polynomial c[0..n-1] (highest degree first,
// n <= 5).  Returns the count or -1 ("unable to find the critical points").
// Warp-uniform call; for degree 3/4 lane i < degree iterates root i.
__device__ __noinline__ int poly_roots_real(const double* c, int n, double* out, int lane) {
  int lead = 0;
  while (lead + 1 < n && c[lead] == 0.0) ++lead;   // RemoveLeadingZeros
  const double* p = c + lead;
  const int degree = n - lead - 1;
  if (degree <= 0) return 0;
  if (degree == 1) {
    out[0] = -p[1] / p[0];
    return 1;
  }
  if (degree == 2) {  // FindQuadraticPolynomialRoots
    const double a = p[0], b = p[1], cc = p[2];
    const double D = b * b - 4 * a * cc;
    const double sq = sqrt(fabs(D));
    if (D >= 0) {
      if (b >= 0) {
        out[0] = (-b - sq) / (2.0 * a);
        out[1] = (2.0 * cc) / (-b - sq);
      } else {
        out[0] = (2.0 * cc) / (-b + sq);
        out[1] = (-b + sq) / (2.0 * a);
      }
    } else {
      out[0] = out[1] = -b / (2.0 * a);
    }
    return 2;
  }
  double m[5];
  double bound = 0.0;
  for (int i = 0; i <= degree; ++i) {
    m[i] = p[i] / p[0];
    if (!isfinite(m[i])) return -1;
    if (i) bound = fmax(bound, fabs(m[i]));
  }
  bound = 0.5 * (bound + 1.0);
  const double kCos3[3] = {0.9210609940028851, -0.79777667414035813, -0.12328431986252686};
  const double kSin3[3] = {0.38941834230865052, 0.60295304808712002, -0.99237139039577016};
  const double kCos4[4] = {0.9210609940028851, -0.38941834230865036, -0.92106099400288521, 0.38941834230865063};
  const double kSin4[4] = {0.38941834230865052, 0.9210609940028851, -0.3894183423086503, -0.92106099400288499};
  const int me = lane < degree ? lane : 0;
  double zr = bound * (degree == 3 ? kCos3[me % 3] : kCos4[me]);
  double zi = bound * (degree == 3 ? kSin3[me % 3] : kSin4[me]);
  for (int it = 0; it < 48; ++it) {
    // Horner for p and p' at this lane's root
    double pr = m[0], pi = 0.0, dr = 0.0, di = 0.0;
    for (int k = 1; k <= degree; ++k) {
      const double ndr = dr * zr - di * zi + pr;
      const double ndi = dr * zi + di * zr + pi;
      dr = ndr;
      di = ndi;
      const double npr = pr * zr - pi * zi + m[k];
      const double npi = pr * zi + pi * zr;
      pr = npr;
      pi = npi;
    }
    double rr = 0.0, ri = 0.0;  // sum_j 1/(z - z_j)
    for (int j = 0; j < degree; ++j) {
      const double ojr = __shfl_sync(0xffffffffu, zr, j), oji = __shfl_sync(0xffffffffu, zi, j);
      if (j == me) continue;
      const double ar = zr - ojr, ai = zi - oji;
      const double inv = 1.0 / (ar * ar + ai * ai);
      rr += ar * inv;
      ri -= ai * inv;
    }
    double wr = 0.0, wi = 0.0;
    if (!(pr == 0.0 && pi == 0.0)) {
      const double inv = 1.0 / (dr * dr + di * di);
      const double nr = (pr * dr + pi * di) * inv, ni = (pi * dr - pr * di) * inv;  // Newton step p/p'
      const double qr = 1.0 - (nr * rr - ni * ri), qi = -(nr * ri + ni * rr);
      const double inv2 = 1.0 / (qr * qr + qi * qi);
      wr = (nr * qr + ni * qi) * inv2;
      wi = (ni * qr - nr * qi) * inv2;
    }
    zr -= wr;
    zi -= wi;
    double change = (wr * wr + wi * wi) / fmax(1e-300, zr * zr + zi * zi);
    if (lane >= degree) change = 0.0;
    for (int o = 4; o > 0; o >>= 1) change = fmax(change, __shfl_xor_sync(0xffffffffu, change, o));
    change = __shfl_sync(0xffffffffu, change, 0);
    if (!(change >= 1e-26)) break;   // |dz| < 1e-13 |z| for every root (or NaN)
  }
  bool bad = false;
  for (int i = 0; i < degree; ++i) {
    out[i] = __shfl_sync(0xffffffffu, zr, i);
    bad = bad || !isfinite(out[i]);
  }
  return bad ? -1 : degree;
}

// Minimiser over [lo, hi] (in x) of the Hermite interpolant through (0, f0, g0),
// (x1, f1, g1) [and (x2, f2, g2) if three == true].  Warp-uniform.
__device__ __noinline__ double hermite_minimizer(double f0, double g0, double x1, double f1, double g1,
                                                 bool three, double x2, double f2, double g2, double lo,
                                                 double hi, int lane) {
  const double h = three ? fmax(x1, x2) : x1;
  const double g0h = g0 * h;
  double c[6];
  int nc;
  if (!three) {
    const double u = f1 - f0 - g0h, v = (g1 - g0) * h;
    c[0] = v - 2.0 * u;
    c[1] = 3.0 * u - v;
    c[2] = g0h;
    c[3] = f0;
    nc = 4;
  } else {
    double A[4][5];
    const double ts[2] = {x1 / h, x2 / h};
    const double fs[2] = {f1, f2}, gs[2] = {g1, g2};
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const double t = ts[q], t2 = t * t, t3 = t2 * t, t4 = t3 * t, t5 = t4 * t;
      A[2 * q][0] = t2; A[2 * q][1] = t3; A[2 * q][2] = t4; A[2 * q][3] = t5;
      A[2 * q][4] = fs[q] - f0 - g0h * t;
      A[2 * q + 1][0] = 2.0 * t; A[2 * q + 1][1] = 3.0 * t2; A[2 * q + 1][2] = 4.0 * t3; A[2 * q + 1][3] = 5.0 * t4;
      A[2 * q + 1][4] = (gs[q] - g0) * h;
    }
    // Gaussian elimination with partial pivoting, fully unrolled (registers)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int i = k + 1; i < 4; ++i) {
        if (fabs(A[i][k]) > fabs(A[k][k])) {
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            const double t = A[k][j];
            A[k][j] = A[i][j];
            A[i][j] = t;
          }
        }
      }
#pragma unroll
      for (int i = k + 1; i < 4; ++i) {
        const double mlt = A[i][k] / A[k][k];
#pragma unroll
        for (int j = k; j < 5; ++j) A[i][j] -= mlt * A[k][j];
      }
    }
    double d[4];
#pragma unroll
    for (int i = 3; i >= 0; --i) {
      double acc = A[i][4];
#pragma unroll
      for (int j = i + 1; j < 4; ++j) acc -= A[i][j] * d[j];
      d[i] = acc / A[i][i];
    }
    c[0] = d[3]; c[1] = d[2]; c[2] = d[1]; c[3] = d[0];
    c[4] = g0h;
    c[5] = f0;
    nc = 6;
  }
  const double tlo = lo / h, thi = hi / h;
  // MinimizePolynomial: middle, ends, real parts of the roots of p'
  double ox = (lo + hi) / 2.0;
  double ov = poly_eval(c, nc, ox / h);
  double v = poly_eval(c, nc, tlo);
  if (v < ov) { ov = v; ox = lo; }
  v = poly_eval(c, nc, thi);
  if (v < ov) { ov = v; ox = hi; }
  double der[5], roots[4];
  const int degree = nc - 1;
  for (int i = 0; i < degree; ++i) der[i] = (degree - i) * c[i];
  const int nr = poly_roots_real(der, degree, roots, lane);
  for (int i = 0; i < nr; ++i) {
    if (roots[i] < tlo || roots[i] > thi) continue;
    v = poly_eval(c, nc, roots[i]);
    if (v < ov) { ov = v; ox = roots[i] * h; }
  }
  // MinimizeInterpolatingPolynomial: the samples themselves, in order
  const double sx[3] = {0.0, x1, x2};
  for (int i = 0; i < (three ? 3 : 2); ++i) {
    if (sx[i] < lo || sx[i] > hi) continue;
    v = poly_eval(c, nc, sx[i] / h);
    if (v < ov) { ov = v; ox = sx[i]; }
  }
  return ox;
}

// One contraction of ArmijoLineSearch::DoSearch: new trial step from
// {initial, current [, previous]}.
__device__ __forceinline__ double ls_next_step(const LsSample& initial, const LsSample& previous,
                                               const LsSample& current, const DevConsts& K, int lane) {
  const double min_step = K.ls_max_contraction * current.x;
  const double max_step = K.ls_min_contraction * current.x;
  // invalid value (or a non-finite gradient, which only overflow can produce):
  // Ceres' bisection rule
  if (!current.value_valid || !current.gradient_valid ||
      (previous.value_valid && !previous.gradient_valid))
    return fmin(fmax(current.x * 0.5, min_step), max_step);
  return hermite_minimizer(initial.value, initial.gradient, current.x, current.value, current.gradient,
                           previous.value_valid, previous.x, previous.value, previous.gradient, min_step,
                           max_step, lane);
}

// warp-wide helpers --------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

}  // namespace lfr
