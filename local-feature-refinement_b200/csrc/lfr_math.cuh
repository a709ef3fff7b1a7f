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
// Line-search step selection (rare path; kept out of line).
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

// Real parts of the roots of p (n coefficients, highest degree first); returns
// the count or -1.  Degree <= 2: closed forms of Ceres polynomial.cc; higher:
// Aberth-Ehrlich iteration (= eigenvalues of the companion matrix).
__device__ __noinline__ int poly_roots_real(const double* pin, int nin, double* out) {
  int lead = 0;
  while (lead + 1 < nin && pin[lead] == 0.0) ++lead;
  const double* p = pin + lead;
  const int degree = nin - lead - 1;
  if (degree <= 0) return 0;
  if (degree == 1) {
    out[0] = -p[1] / p[0];
    return 1;
  }
  if (degree == 2) {
    const double a = p[0], b = p[1], c = p[2];
    const double D = b * b - 4 * a * c;
    const double sq = sqrt(fabs(D));
    if (D >= 0) {
      if (b >= 0) {
        out[0] = (-b - sq) / (2.0 * a);
        out[1] = (2.0 * c) / (-b - sq);
      } else {
        out[0] = (2.0 * c) / (-b + sq);
        out[1] = (-b + sq) / (2.0 * a);
      }
    } else {
      out[0] = out[1] = -b / (2.0 * a);
    }
    return 2;
  }
  if (degree > 5) return -1;
  double m[6];
  for (int i = 0; i <= degree; ++i) {
    m[i] = p[i] / p[0];
    if (!isfinite(m[i])) return -1;
  }
  double bound = 0.0;
  for (int i = 1; i <= degree; ++i) bound = fmax(bound, fabs(m[i]));
  bound += 1.0;
  double zr[5], zi[5];
  for (int i = 0; i < degree; ++i) {
    const double ang = 2.0 * 3.14159265358979323846 * i / degree + 0.4;
    zr[i] = 0.5 * bound * cos(ang);
    zi[i] = 0.5 * bound * sin(ang);
  }
  for (int it = 0; it < 64; ++it) {
    double change = 0.0;
    for (int i = 0; i < degree; ++i) {
      // Horner for p and p' at z_i (complex)
      double pr = m[0], pi = 0.0, dr = 0.0, di = 0.0;
      for (int k = 1; k <= degree; ++k) {
        const double ndr = dr * zr[i] - di * zi[i] + pr;
        const double ndi = dr * zi[i] + di * zr[i] + pi;
        dr = ndr;
        di = ndi;
        const double npr = pr * zr[i] - pi * zi[i] + m[k];
        const double npi = pr * zi[i] + pi * zr[i];
        pr = npr;
        pi = npi;
      }
      if (pr == 0.0 && pi == 0.0) continue;
      // newton = p / p'
      double den = dr * dr + di * di;
      const double nr = (pr * dr + pi * di) / den, ni = (pi * dr - pr * di) / den;
      double rr = 0.0, ri = 0.0;  // sum 1/(z_i - z_j)
      for (int j = 0; j < degree; ++j) {
        if (j == i) continue;
        const double ar = zr[i] - zr[j], ai = zi[i] - zi[j];
        const double d2 = ar * ar + ai * ai;
        rr += ar / d2;
        ri -= ai / d2;
      }
      // w = newton / (1 - newton * repel)
      const double qr = 1.0 - (nr * rr - ni * ri), qi = -(nr * ri + ni * rr);
      den = qr * qr + qi * qi;
      const double wr = (nr * qr + ni * qi) / den, wi = (ni * qr - nr * qi) / den;
      zr[i] -= wr;
      zi[i] -= wi;
      change = fmax(change, sqrt(wr * wr + wi * wi) / fmax(1e-300, sqrt(zr[i] * zr[i] + zi[i] * zi[i])));
    }
    if (change < 1e-13) break;  // roots to 1e-13 relative: far below what the step choice resolves
  }
  for (int i = 0; i < degree; ++i) {
    if (!isfinite(zr[i])) return -1;
    out[i] = zr[i];
  }
  return degree;
}

// MinimizeInterpolatingPolynomial over [x_min, x_max] for up to 3 samples
// (value + gradient each => polynomial degree <= 5).
__device__ __noinline__ double ls_minimize_interpolant(const LsSample* s, int ns, double x_min,
                                                        double x_max) {
  int nc = 0;
  for (int i = 0; i < ns; ++i) nc += (s[i].value_valid ? 1 : 0) + (s[i].gradient_valid ? 1 : 0);
  const int degree = nc - 1;
  double A[36], b[6], poly[6];
  int perm[6];
  for (int i = 0; i < 36; ++i) A[i] = 0.0;
  int row = 0;
  for (int i = 0; i < ns; ++i) {
    if (s[i].value_valid) {
      for (int j = 0; j <= degree; ++j) A[row * nc + j] = pow(s[i].x, (double)(degree - j));
      b[row++] = s[i].value;
    }
    if (s[i].gradient_valid) {
      for (int j = 0; j < degree; ++j)
        A[row * nc + j] = (degree - j) * pow(s[i].x, (double)(degree - j - 1));
      b[row++] = s[i].gradient;
    }
  }
  // full-pivot Gaussian elimination (Eigen fullPivLu().solve)
  for (int i = 0; i < nc; ++i) perm[i] = i;
  for (int k = 0; k < nc; ++k) {
    int pr = k, pc = k;
    double best = -1.0;
    for (int i = k; i < nc; ++i)
      for (int j = k; j < nc; ++j)
        if (fabs(A[i * nc + j]) > best) {
          best = fabs(A[i * nc + j]);
          pr = i;
          pc = j;
        }
    if (best == 0.0) break;
    if (pr != k) {
      for (int j = 0; j < nc; ++j) {
        const double t = A[k * nc + j];
        A[k * nc + j] = A[pr * nc + j];
        A[pr * nc + j] = t;
      }
      const double t = b[k];
      b[k] = b[pr];
      b[pr] = t;
    }
    if (pc != k) {
      for (int i = 0; i < nc; ++i) {
        const double t = A[i * nc + k];
        A[i * nc + k] = A[i * nc + pc];
        A[i * nc + pc] = t;
      }
      const int t = perm[k];
      perm[k] = perm[pc];
      perm[pc] = t;
    }
    for (int i = k + 1; i < nc; ++i) {
      const double mlt = A[i * nc + k] / A[k * nc + k];
      if (mlt == 0.0) continue;
      for (int j = k; j < nc; ++j) A[i * nc + j] -= mlt * A[k * nc + j];
      b[i] -= mlt * b[k];
    }
  }
  double yv[6];
  for (int i = nc - 1; i >= 0; --i) {
    double acc = b[i];
    for (int j = i + 1; j < nc; ++j) acc -= A[i * nc + j] * yv[j];
    yv[i] = (A[i * nc + i] != 0.0) ? acc / A[i * nc + i] : 0.0;
  }
  for (int i = 0; i < nc; ++i) poly[perm[i]] = yv[i];

  // MinimizePolynomial: middle, ends, then real parts of the roots of p'
  double ox = (x_min + x_max) / 2.0;
  double ov = poly_eval(poly, nc, ox);
  double v = poly_eval(poly, nc, x_min);
  if (v < ov) { ov = v; ox = x_min; }
  v = poly_eval(poly, nc, x_max);
  if (v < ov) { ov = v; ox = x_max; }
  if (nc > 2) {
    double der[5], roots[5];
    for (int i = 0; i < degree; ++i) der[i] = (degree - i) * poly[i];
    const int nr = poly_roots_real(der, degree, roots);
    for (int i = 0; i < nr; ++i) {
      if (roots[i] < x_min || roots[i] > x_max) continue;
      v = poly_eval(poly, nc, roots[i]);
      if (v < ov) { ov = v; ox = roots[i]; }
    }
  }
  for (int i = 0; i < ns; ++i) {
    if (s[i].x < x_min || s[i].x > x_max) continue;
    v = poly_eval(poly, nc, s[i].x);
    if (v < ov) { ov = v; ox = s[i].x; }
  }
  return ox;
}

// One contraction of ArmijoLineSearch::DoSearch: new trial step from
// {initial, current [, previous]}.
__device__ __noinline__ double ls_next_step(const LsSample& initial, const LsSample& previous,
                                            const LsSample& current, const DevConsts& K) {
  const double min_step = K.ls_max_contraction * current.x;
  const double max_step = K.ls_min_contraction * current.x;
  if (!current.value_valid) return fmin(fmax(current.x * 0.5, min_step), max_step);
  LsSample s[3];
  s[0] = initial;
  s[1] = current;
  int ns = 2;
  if (previous.value_valid) s[ns++] = previous;
  return ls_minimize_interpolant(s, ns, min_step, max_step);
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
