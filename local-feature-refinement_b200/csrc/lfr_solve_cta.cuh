// lfr_solve_cta.cuh — one CTA solves one LARGE component (more than 96
// unknowns: Madrid-scale scenes cap a component at #images = 1000 nodes,
// solve.cc:586), or every component when lfr_options.linear_solver = 2.
//
// Same trust-region loop as the warp kernels, but the
// damped normal equations (S H S + D^2) y = S g are never formed: they are
// solved matrix-free by conjugate gradients preconditioned with the inverse of
// the 2x2 diagonal blocks (block-Jacobi), iterated to a relative residual of
// 1e-14 and then refined on the true residual to 2e-15, so that the step is, to working precision, the exact solve Ceres'
// SPARSE_NORMAL_CHOLESKY (solve.cc:147) returns:
//
//   q_e = a_e (x_d~ - M_e x_s~)             one thread per directed edge
//   w_v = S_v ( sum_out -M_e^T q_e + sum_in q_e ) + D_v^2 p_v   one thread per free node
//
// with x~ = S p, a_e = sim rho', M_e = I + grad(flow) staged by the evaluation.
// The kept-edge lists, in-edge lists and free-variable numbering of these few
// large components are prepared on the host (lfr_capi.cu) and live in HBM.
#pragma once
#include "lfr_solve_warp.cuh"

namespace lfr {

struct CtaComp {
  uint32_t slot, Nc, Ec, nf;
  uint64_t e_off, n_off, f_off;  // offsets of this component in the per-edge / per-node / per-free-node arrays
  uint32_t comp_index, regular;  // ordinal among the large components; 1 = every kept edge has exactly one twin
};

struct CtaArrays {
  // per kept edge
  const uint32_t* eidx;    // global edge index
  const uint32_t* meta;    // src_local | dst_local << 14 | kind << 28
  const uint32_t* inlist;  // kept-edge indices sorted by destination
  const uint32_t* twin;    // index of the reverse edge (regular components)
  const int32_t* fdst;     // free index of the edge's destination node, or -1
  double* bmat;            // 4 doubles per edge: scaled off-diagonal block S_s H_sd S_d (regular components)
  double* scr;             // 7 doubles per edge (a, r0, r1, m00, m01, m10, m11), SoA per component
  double* q;               // 2 doubles per edge (matvec scratch)
  // per node
  const uint32_t* node;    // global node index
  const uint32_t* outptr;  // [Nc + 1] per component
  const uint32_t* inptr;   // [Nc + 1] per component
  const int32_t* freeof;   // free index or -1
  double* x;               // 2 per node
  double* xc;              // 2 per node
  // per free node
  const uint32_t* lof;     // local node of free index
  double* vec;             // 13 vectors of 2 doubles per free node + 2 of 3 doubles, see offsets below
  uint64_t total_free;     // sum of nf over all large components (stride between vectors)
};

constexpr int kCtaThreads = 256;
enum { V_G = 0, V_S, V_DL, V_D2, V_R, V_Z, V_P, V_W, V_Y, V_COUNT };  // 2-doubles-per-node vectors
// then three 3-doubles-per-node arrays: diagonal blocks (d00, d01, d11), their damped scaled form, its inverse

struct CtaCtx {
  int tid, Nc, Ec, nf, n;
  const uint32_t *eidx, *meta, *inlist, *node, *outptr, *inptr, *lof;
  const int32_t* freeof;
  double *scr, *q, *x, *xc, *g, *S, *dl, *D2, *r, *z, *p, *w, *y, *diag, *pinv, *pblk, *bmat;
  const uint32_t* twin;
  const int32_t* fdst;
  bool regular;
  const float4* edges;
  double* red;  // shared: 3 * 8 doubles
};

__device__ __forceinline__ void block_sum3(const CtaCtx& C, double& a, double& b, double& c) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(kFull, a, o);
    b += __shfl_xor_sync(kFull, b, o);
    c += __shfl_xor_sync(kFull, c, o);
  }
  const int w = C.tid >> 5, nw = kCtaThreads / 32;
  __syncthreads();  // previous users of `red` are done
  if ((C.tid & 31) == 0) {
    C.red[w] = a;
    C.red[nw + w] = b;
    C.red[2 * nw + w] = c;
  }
  __syncthreads();
  a = b = c = 0.0;
#pragma unroll
  for (int i = 0; i < nw; ++i) {  // fixed order: identical in every thread, reproducible
    a += C.red[i];
    b += C.red[nw + i];
    c += C.red[2 * nw + i];
  }
}
__device__ __forceinline__ double block_max(const CtaCtx& C, double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  const int w = C.tid >> 5, nw = kCtaThreads / 32;
  __syncthreads();
  if ((C.tid & 31) == 0) C.red[w] = v;
  __syncthreads();
  v = C.red[0];
#pragma unroll
  for (int i = 1; i < nw; ++i) v = fmax(v, C.red[i]);
  return v;
}

__device__ __forceinline__ double cta_eval(const CtaCtx& C, const double* xe, const DevConsts& K) {
  double cost = 0.0, z1 = 0.0, z2 = 0.0;
  const int E = C.Ec;
  for (int j = C.tid; j < E; j += kCtaThreads) {
    const uint32_t mt = C.meta[j];
    const int s = mt & 0x3fff, d = (mt >> 14) & 0x3fff, kind = mt >> 28;
    const float4* qp = C.edges + 5 * (size_t)C.eidx[j];
    float4 q[5];
#pragma unroll
    for (int t = 0; t < 5; ++t) q[t] = __ldg(qp + t);
    const EdgeEval ev = eval_edge(q, kind, xe[2 * s], xe[2 * s + 1], xe[2 * d], xe[2 * d + 1], K);
    double* sc = C.scr + j;
    sc[0] = ev.a;
    sc[E] = ev.r0;
    sc[2 * (size_t)E] = ev.r1;
    sc[3 * (size_t)E] = ev.m00;
    sc[4 * (size_t)E] = ev.m01;
    sc[5 * (size_t)E] = ev.m10;
    sc[6 * (size_t)E] = ev.m11;
    cost += ev.half_rho;
  }
  block_sum3(C, cost, z1, z2);
  return cost;
}

// Diagonal blocks + gradient from the staged evaluation (GRAD_ONLY: returns grad . dl).
template <bool GRAD_ONLY>
__device__ __forceinline__ double cta_assemble(const CtaCtx& C, bool first, const DevConsts& K) {
  const size_t E = C.Ec;
  double acc = 0.0, gmax = 0.0;
  for (int f = C.tid; f < C.nf; f += kCtaThreads) {
    const int l = C.lof[f];
    double d00 = 0., d01 = 0., d11 = 0., g0 = 0., g1 = 0.;
    for (uint32_t j = C.outptr[l]; j < C.outptr[l + 1]; ++j) {
      const double a = C.scr[j], r0 = C.scr[E + j], r1 = C.scr[2 * E + j];
      const double m00 = C.scr[3 * E + j], m01 = C.scr[4 * E + j], m10 = C.scr[5 * E + j], m11 = C.scr[6 * E + j];
      g0 -= a * (m00 * r0 + m10 * r1);
      g1 -= a * (m01 * r0 + m11 * r1);
      if (!GRAD_ONLY) {
        d00 += a * (m00 * m00 + m10 * m10);
        d01 += a * (m00 * m01 + m10 * m11);
        d11 += a * (m01 * m01 + m11 * m11);
      }
    }
    for (uint32_t t = C.inptr[l]; t < C.inptr[l + 1]; ++t) {
      const uint32_t j = C.inlist[t];
      const double a = C.scr[j];
      g0 += a * C.scr[E + j];
      g1 += a * C.scr[2 * E + j];
      if (!GRAD_ONLY) {
        d00 += a;
        d11 += a;
      }
    }
    if (GRAD_ONLY) {
      acc += g0 * C.dl[2 * f] + g1 * C.dl[2 * f + 1];
    } else {
      C.diag[3 * f] = d00;
      C.diag[3 * f + 1] = d01;
      C.diag[3 * f + 2] = d11;
      C.g[2 * f] = g0;
      C.g[2 * f + 1] = g1;
      if (first) {
        C.S[2 * f] = 1.0 / (1.0 + sqrt(d00));
        C.S[2 * f + 1] = 1.0 / (1.0 + sqrt(d11));
      }
      const double x0 = C.x[2 * l], x1 = C.x[2 * l + 1];
      const double p0 = fmin(fmax(x0 - g0, -K.bound), K.bound), p1 = fmin(fmax(x1 - g1, -K.bound), K.bound);
      gmax = fmax(gmax, fmax(fabs(x0 - p0), fabs(x1 - p1)));
    }
  }
  if (GRAD_ONLY) {
    double z1 = 0.0, z2 = 0.0;
    block_sum3(C, acc, z1, z2);
    return acc;
  }
  gmax = block_max(C, gmax);  // (also makes S visible to every thread)
  if (C.regular) {
    // block-CSR form of S H S: one scaled 2x2 block per out-edge, edge + twin combined
    for (int j = C.tid; j < C.Ec; j += kCtaThreads) {
      const uint32_t mt = C.meta[j];
      const int fs = C.freeof[mt & 0x3fff], fd = C.fdst[j];
      if (fs < 0 || fd < 0) continue;
      const uint32_t t = C.twin[j];
      const double a = C.scr[j], at = C.scr[t];
      const double m00 = C.scr[3 * E + j], m01 = C.scr[4 * E + j], m10 = C.scr[5 * E + j], m11 = C.scr[6 * E + j];
      const double t00 = C.scr[3 * E + t], t01 = C.scr[4 * E + t], t10 = C.scr[5 * E + t], t11 = C.scr[6 * E + t];
      const double s0 = C.S[2 * fs], s1 = C.S[2 * fs + 1], d0 = C.S[2 * fd], d1 = C.S[2 * fd + 1];
      double* b = C.bmat + 4 * (size_t)j;  // block (s, d) = -a M^T - a_t M_t
      b[0] = s0 * (-a * m00 - at * t00) * d0;
      b[1] = s0 * (-a * m10 - at * t01) * d1;
      b[2] = s1 * (-a * m01 - at * t10) * d0;
      b[3] = s1 * (-a * m11 - at * t11) * d1;
    }
    __syncthreads();
  }
  return gmax;
}

// w = (S H S + D^2) v, matrix-free.
__device__ __forceinline__ void cta_matvec(const CtaCtx& C, const double* v, double* w) {
  const size_t E = C.Ec;
#pragma unroll 4
  for (int j = C.tid; j < C.Ec; j += kCtaThreads) {
    const uint32_t mt = C.meta[j];
    const int fs = C.freeof[mt & 0x3fff], fd = C.freeof[(mt >> 14) & 0x3fff];
    double s0 = 0., s1 = 0., t0 = 0., t1 = 0.;
    if (fs >= 0) {
      s0 = C.S[2 * fs] * v[2 * fs];
      s1 = C.S[2 * fs + 1] * v[2 * fs + 1];
    }
    if (fd >= 0) {
      t0 = C.S[2 * fd] * v[2 * fd];
      t1 = C.S[2 * fd + 1] * v[2 * fd + 1];
    }
    const double a = C.scr[j];
    const double m00 = C.scr[3 * E + j], m01 = C.scr[4 * E + j], m10 = C.scr[5 * E + j], m11 = C.scr[6 * E + j];
    C.q[2 * (size_t)j] = a * (t0 - (m00 * s0 + m01 * s1));      // a (J x~)_0,  J = [-M | I]
    C.q[2 * (size_t)j + 1] = a * (t1 - (m10 * s0 + m11 * s1));
  }
  __syncthreads();
  for (int f = C.tid; f < C.nf; f += kCtaThreads) {
    const int l = C.lof[f];
    double a0 = 0., a1 = 0.;
#pragma unroll 4
    for (uint32_t j = C.outptr[l]; j < C.outptr[l + 1]; ++j) {
      const double q0 = C.q[2 * (size_t)j], q1 = C.q[2 * (size_t)j + 1];
      a0 -= C.scr[3 * E + j] * q0 + C.scr[5 * E + j] * q1;   // -M^T q
      a1 -= C.scr[4 * E + j] * q0 + C.scr[6 * E + j] * q1;
    }
#pragma unroll 4
    for (uint32_t t = C.inptr[l]; t < C.inptr[l + 1]; ++t) {
      const uint32_t j = C.inlist[t];
      a0 += C.q[2 * (size_t)j];
      a1 += C.q[2 * (size_t)j + 1];
    }
    w[2 * f] = C.S[2 * f] * a0 + C.D2[2 * f] * v[2 * f];
    w[2 * f + 1] = C.S[2 * f + 1] * a1 + C.D2[2 * f + 1] * v[2 * f + 1];
  }
  __syncthreads();
}

// Regular components: w = (S H S + D^2) v in ONE node-parallel pass over the
// block-CSR rows (damped diagonal block `pblk` + one 2x2 block per out-edge);
// returns this thread's share of v . w.
__device__ __forceinline__ double cta_matvec_bcsr(const CtaCtx& C, const double* v, double* w, const double* pblk) {
  double dot = 0.0;
  for (int f = C.tid; f < C.nf; f += kCtaThreads) {
    const int l = C.lof[f];
    const double v0 = v[2 * f], v1 = v[2 * f + 1];
    double a0 = pblk[3 * f] * v0 + pblk[3 * f + 1] * v1, a1 = pblk[3 * f + 1] * v0 + pblk[3 * f + 2] * v1;
#pragma unroll 4
    for (uint32_t j = C.outptr[l]; j < C.outptr[l + 1]; ++j) {
      const int fd = C.fdst[j];
      if (fd < 0) continue;
      const double* b = C.bmat + 4 * (size_t)j;
      const double u0 = v[2 * fd], u1 = v[2 * fd + 1];
      a0 += b[0] * u0 + b[1] * u1;
      a1 += b[2] * u0 + b[3] * u1;
    }
    w[2 * f] = a0;
    w[2 * f + 1] = a1;
    dot += v0 * a0 + v1 * a1;
  }
  __syncthreads();
  return dot;
}

// Block-Jacobi PCG for (S H S + D^2) y = S g; dl = -S y.  Returns validity and
// {model_cost_change, g . dl, |dl|_inf}.
__device__ __forceinline__ bool cta_lm_step(const CtaCtx& C, double radius, const DevConsts& K, double* model_change,
                                            double* gd, double* dmax, unsigned* cg_iters, bool diag_mode,
                                            double* worst_res, unsigned* n_maxit) {
  // damping, preconditioner, initial residual
  double bb = 0.0, rz = 0.0, bad = 0.0;
  for (int f = C.tid; f < C.nf; f += kCtaThreads) {
    const double s0 = C.S[2 * f], s1 = C.S[2 * f + 1];
    const double h00 = C.diag[3 * f] * s0 * s0, h01 = C.diag[3 * f + 1] * s0 * s1, h11 = C.diag[3 * f + 2] * s1 * s1;
    const double e0 = fmin(fmax(h00, K.min_diag), K.max_diag) / radius;
    const double e1 = fmin(fmax(h11, K.min_diag), K.max_diag) / radius;
    C.D2[2 * f] = e0;
    C.D2[2 * f + 1] = e1;
    const double p00 = h00 + e0, p11 = h11 + e1;
    const double det = p00 * p11 - h01 * h01;
    if (!(det > 0.0) || !(p00 > 0.0)) bad = 1.0;
    const double id = 1.0 / det;
    C.pblk[3 * f] = p00;
    C.pblk[3 * f + 1] = h01;
    C.pblk[3 * f + 2] = p11;
    C.pinv[3 * f] = p11 * id;
    C.pinv[3 * f + 1] = -h01 * id;
    C.pinv[3 * f + 2] = p00 * id;
    const double b0 = s0 * C.g[2 * f], b1 = s1 * C.g[2 * f + 1];
    C.y[2 * f] = 0.0;
    C.y[2 * f + 1] = 0.0;
    C.r[2 * f] = b0;
    C.r[2 * f + 1] = b1;
    const double z0 = C.pinv[3 * f] * b0 + C.pinv[3 * f + 1] * b1, z1 = C.pinv[3 * f + 1] * b0 + C.pinv[3 * f + 2] * b1;
    C.z[2 * f] = z0;
    C.z[2 * f + 1] = z1;
    C.p[2 * f] = z0;
    C.p[2 * f + 1] = z1;
    bb += b0 * b0 + b1 * b1;
    rz += b0 * z0 + b1 * z1;
  }
  block_sum3(C, bb, rz, bad);
  bool ok = (bad == 0.0) && isfinite(bb);
  const double tol2 = 1e-28 * bb;  // recurrence residual |r| <= 1e-14 |b| ...
  const int max_it = 8 * C.n + 100;
  int it = 0, restarts = 0;
  // ... then iterative refinement: the recurrence residual drifts away from b - A y by ~cond * eps,
  // so CG is restarted from the TRUE residual until that is <= 2e-15 |b| (at most twice): the step is
  // the exact solve of Ceres' SPARSE_NORMAL_CHOLESKY to working precision, which keeps the line
  // search's discontinuous decisions on the oracle's side
  while (ok && bb > 0.0) {
    for (; it < max_it; ++it) {
      double pw = 0.0, z1 = 0.0, z2 = 0.0;
      if (C.regular) {
        pw = cta_matvec_bcsr(C, C.p, C.w, C.pblk);
      } else {
        cta_matvec(C, C.p, C.w);
        for (int i = C.tid; i < C.n; i += kCtaThreads) pw += C.p[i] * C.w[i];
      }
      block_sum3(C, pw, z1, z2);
      if (!(pw > 0.0) || !isfinite(pw)) {  // not positive definite in working precision
        ok = false;
        break;
      }
      const double alpha = rz / pw;
      double rr = 0.0, rz_new = 0.0, z3 = 0.0;
      for (int f = C.tid; f < C.nf; f += kCtaThreads) {
        const double y0 = C.y[2 * f] + alpha * C.p[2 * f], y1 = C.y[2 * f + 1] + alpha * C.p[2 * f + 1];
        const double r0 = C.r[2 * f] - alpha * C.w[2 * f], r1 = C.r[2 * f + 1] - alpha * C.w[2 * f + 1];
        C.y[2 * f] = y0;
        C.y[2 * f + 1] = y1;
        C.r[2 * f] = r0;
        C.r[2 * f + 1] = r1;
        const double z0 = C.pinv[3 * f] * r0 + C.pinv[3 * f + 1] * r1, zz1 = C.pinv[3 * f + 1] * r0 + C.pinv[3 * f + 2] * r1;
        C.z[2 * f] = z0;
        C.z[2 * f + 1] = zz1;
        rr += r0 * r0 + r1 * r1;
        rz_new += r0 * z0 + r1 * zz1;
      }
      block_sum3(C, rr, rz_new, z3);
      if (rr <= tol2) {
        ++it;
        break;
      }
      const double beta = rz_new / rz;
      rz = rz_new;
      for (int i = C.tid; i < C.n; i += kCtaThreads) C.p[i] = C.z[i] + beta * C.p[i];
      __syncthreads();
    }
    if (!ok || it >= max_it || restarts >= 2) break;
    // true residual r = S g - A y, z = M^-1 r, p = z
    if (C.regular) cta_matvec_bcsr(C, C.y, C.w, C.pblk); else cta_matvec(C, C.y, C.w);
    double rt = 0.0, rzt = 0.0, z3 = 0.0;
    for (int f = C.tid; f < C.nf; f += kCtaThreads) {
      const double r0 = C.S[2 * f] * C.g[2 * f] - C.w[2 * f], r1 = C.S[2 * f + 1] * C.g[2 * f + 1] - C.w[2 * f + 1];
      const double z0 = C.pinv[3 * f] * r0 + C.pinv[3 * f + 1] * r1, z1 = C.pinv[3 * f + 1] * r0 + C.pinv[3 * f + 2] * r1;
      C.r[2 * f] = r0;
      C.r[2 * f + 1] = r1;
      C.z[2 * f] = z0;
      C.z[2 * f + 1] = z1;
      C.p[2 * f] = z0;
      C.p[2 * f + 1] = z1;
      rt += r0 * r0 + r1 * r1;
      rzt += r0 * z0 + r1 * z1;
    }
    block_sum3(C, rt, rzt, z3);
    if (rt <= 4e-30 * bb || !(rzt > 0.0)) break;  // |b - A y| <= 2e-15 |b|
    rz = rzt;
    ++restarts;
  }
  *cg_iters += (unsigned)it;
  if (diag_mode && ok && bb > 0.0) {  // LFR_PROFILE: true residual |b - A y| / |b| of the returned solution
    if (it >= max_it) ++*n_maxit;
    if (C.regular) cta_matvec_bcsr(C, C.y, C.w, C.pblk); else cta_matvec(C, C.y, C.w);
    double e2 = 0.0, z1 = 0.0, z2 = 0.0;
    for (int i = C.tid; i < C.n; i += kCtaThreads) {
      const double d = C.S[i] * C.g[i] - C.w[i];
      e2 += d * d;
    }
    block_sum3(C, e2, z1, z2);
    *worst_res = fmax(*worst_res, sqrt(e2 / bb));
  }
  double mc = 0.0, dot = 0.0, nonfinite = 0.0, mx = 0.0;
  for (int i = C.tid; i < C.n; i += kCtaThreads) {
    const double y = C.y[i], si = C.S[i], gi = C.g[i];
    const double d = -si * y;
    C.dl[i] = d;
    mc += y * (si * gi + C.D2[i] * y);
    dot += gi * d;
    mx = fmax(mx, fabs(d));
    if (!isfinite(y)) nonfinite = 1.0;
  }
  block_sum3(C, mc, dot, nonfinite);
  mx = block_max(C, mx);
  *model_change = 0.5 * mc;
  *gd = dot;
  *dmax = mx;
  return ok && nonfinite == 0.0;
}

__device__ __forceinline__ void cta_candidate(const CtaCtx& C, double alpha, const DevConsts& K) {
  for (int i = C.tid; i < 2 * C.Nc; i += kCtaThreads) {
    const int f = C.freeof[i >> 1];
    double v = C.x[i];
    if (f >= 0) v = fmin(fmax(v + alpha * C.dl[2 * f + (i & 1)], -K.bound), K.bound);
    C.xc[i] = v;
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCtaThreads)
solve_cta_kernel(const DevProblem P, const DevConsts K, const CtaArrays A, const CtaComp* comps, unsigned smem_doubles) {
  __shared__ double red[3 * (kCtaThreads / 32)];
  const CtaComp cc = comps[blockIdx.x];
  CtaCtx C;
  C.tid = threadIdx.x;
  C.Nc = (int)cc.Nc;
  C.Ec = (int)cc.Ec;
  C.nf = (int)cc.nf;
  C.n = 2 * C.nf;
  C.eidx = A.eidx + cc.e_off;
  C.meta = A.meta + cc.e_off;
  C.inlist = A.inlist + cc.e_off;
  C.scr = A.scr + 7 * cc.e_off;
  C.q = A.q + 2 * cc.e_off;
  C.node = A.node + cc.n_off;
  C.outptr = A.outptr + cc.n_off + cc.comp_index;
  C.inptr = A.inptr + cc.n_off + cc.comp_index;
  C.freeof = A.freeof + cc.n_off;
  C.x = A.x + 2 * cc.n_off;
  C.xc = A.xc + 2 * cc.n_off;
  C.lof = A.lof + cc.f_off;
  const uint64_t stride = 2 * A.total_free;
  double* v0 = A.vec + 2 * cc.f_off;
  C.g = v0 + V_G * stride;
  C.S = v0 + V_S * stride;
  C.dl = v0 + V_DL * stride;
  C.D2 = v0 + V_D2 * stride;
  C.r = v0 + V_R * stride;
  C.z = v0 + V_Z * stride;
  C.p = v0 + V_P * stride;  // (replaced by shared memory below when it fits)
  C.w = v0 + V_W * stride;
  C.y = v0 + V_Y * stride;
  C.diag = A.vec + V_COUNT * stride + 3 * cc.f_off;
  C.pblk = A.vec + V_COUNT * stride + 3 * A.total_free + 3 * cc.f_off;
  C.pinv = A.vec + V_COUNT * stride + 6 * A.total_free + 3 * cc.f_off;
  C.bmat = A.bmat + 4 * cc.e_off;
  C.twin = A.twin + cc.e_off;
  C.fdst = A.fdst + cc.e_off;
  C.regular = cc.regular != 0;
  // CG vectors on chip while they fit in the launch's dynamic shared memory, in order of how often an
  // iteration touches them: p (randomly gathered by the matvec), w, r, z, y, then the preconditioner
  extern __shared__ __align__(16) double cg_smem[];
  {
    size_t used = 0;
    const size_t cap = smem_doubles;
    auto take = [&](double*& ptr, size_t n_doubles) {
      if (used + n_doubles <= cap) {
        ptr = cg_smem + used;
        used += n_doubles;
      }
    };
    take(C.p, (size_t)C.n);
    take(C.w, (size_t)C.n);
    take(C.r, (size_t)C.n);
    take(C.z, (size_t)C.n);
    take(C.y, (size_t)C.n);
    take(C.pinv, 3 * (size_t)C.nf);
  }
  C.edges = P.edges;
  C.red = red;
  const uint32_t c = cc.slot;
  const int tid = C.tid, lane = tid & 31;

  // start point: IterationZero projects the free blocks onto the box
  for (int l = tid; l < C.Nc; l += kCtaThreads) {
    const uint32_t v = C.node[l];
    double p0 = P.positions[2 * (size_t)v], p1 = P.positions[2 * (size_t)v + 1];
    if (C.freeof[l] >= 0) {
      p0 = fmin(fmax(p0, -K.bound), K.bound);
      p1 = fmin(fmax(p1, -K.bound), K.bound);
    }
    C.x[2 * l] = p0;
    C.x[2 * l + 1] = p1;
  }
  __syncthreads();
  if (tid == 0) P.st_kept[c] = cc.Ec;
  if (C.nf == 0) {
    if (tid == 0) {
      P.st_iter[c] = 0;
      P.st_term[c] = LFR_TERM_EMPTY;
      P.st_cost0[c] = 0.0;
      P.st_cost1[c] = 0.0;
      P.st_ls[c] = 0;
    }
    return;
  }
  double cost = cta_eval(C, C.x, K);
  double gmax = cta_assemble<false>(C, true, K);
  const double cost0 = cost;
  double radius = K.radius0, nu = 2.0;
  int iter = 0, n_invalid = 0, term = LFR_TERM_NO_CONVERGENCE;
  unsigned ls_steps = 0, cg_iters = 0, n_maxit = 0;
  double worst_res = 0.0;
  bool success = true;
  auto norms = [&](double* xn2, double* dn2) {
    double a = 0.0, b = 0.0, z = 0.0;
    for (int i = tid; i < C.n; i += kCtaThreads) {
      const int l = C.lof[i >> 1];
      const double xv = C.x[2 * l + (i & 1)], dv = xv - C.xc[2 * l + (i & 1)];
      a += xv * xv;
      b += dv * dv;
    }
    block_sum3(C, a, b, z);
    *xn2 = a;
    *dn2 = b;
  };
  double x_norm;
  {
    double a = 0.0, b = 0.0, z = 0.0;
    for (int i = tid; i < C.n; i += kCtaThreads) {
      const int l = C.lof[i >> 1];
      const double xv = C.x[2 * l + (i & 1)];
      a += xv * xv;
    }
    block_sum3(C, a, b, z);
    x_norm = sqrt(a);
  }
  for (;;) {
    if (iter >= K.max_iter) { term = LFR_TERM_NO_CONVERGENCE; break; }
    if (success && gmax <= K.g_tol) { term = LFR_TERM_GRADIENT_TOL; break; }
    if (radius <= K.radius_min) { term = LFR_TERM_MIN_RADIUS; break; }
    ++iter;
    success = false;
    double model_change = 0.0, gd = 0.0, dmax = 0.0;
    bool valid = cta_lm_step(C, radius, K, &model_change, &gd, &dmax, &cg_iters, P.st_cycles != nullptr, &worst_res,
                             &n_maxit);
    valid = valid && (model_change > 0.0);
    if (!valid) {
      if (++n_invalid >= K.max_invalid) { term = LFR_TERM_FAILURE; break; }
      radius /= nu;
      nu *= 2.0;
      continue;
    }
    n_invalid = 0;
    cta_candidate(C, 1.0, K);
    double cost_c = cta_eval(C, C.xc, K);
    bool c_valid = isfinite(cost_c);
    if (!c_valid || cost_c > cost + K.ls_suff * gd * 1.0) {
      LsSample initial{0.0, cost, gd, true, true};
      LsSample previous{0.0, 0.0, 0.0, false, false};
      LsSample current{1.0, cost_c, 0.0, c_valid, false};
      if (c_valid) {
        current.gradient = cta_assemble<true>(C, false, K);
        current.gradient_valid = isfinite(current.gradient);
      }
      int ls_iter = 0;
      bool ls_ok = false;
      for (;;) {
        ++ls_iter;
        ++ls_steps;
        if (ls_iter >= K.max_ls_iter) break;
        const double step = ls_next_step(initial, previous, current, K, lane);
        if (step * dmax < K.ls_min_step) break;
        previous = current;
        cta_candidate(C, step, K);
        cost_c = cta_eval(C, C.xc, K);
        c_valid = isfinite(cost_c);
        current = LsSample{step, cost_c, 0.0, c_valid, false};
        if (c_valid) {
          current.gradient = cta_assemble<true>(C, false, K);
          current.gradient_valid = isfinite(current.gradient);
        }
        if (c_valid && !(cost_c > cost + K.ls_suff * gd * step)) { ls_ok = true; break; }
      }
      if (ls_ok) {
        for (int i = tid; i < C.n; i += kCtaThreads) C.dl[i] *= current.x;
        __syncthreads();
      } else {
        cta_candidate(C, 1.0, K);
        cost_c = cta_eval(C, C.xc, K);
        c_valid = isfinite(cost_c);
      }
    }
    if (!c_valid) cost_c = 1.7976931348623157e308;
    double xn2, dn2;
    norms(&xn2, &dn2);
    const double step_norm = sqrt(dn2);
    if (step_norm <= K.p_tol * (x_norm + K.p_tol)) { term = LFR_TERM_PARAMETER_TOL; break; }
    if (fabs(cost - cost_c) <= K.f_tol * cost) { term = LFR_TERM_FUNCTION_TOL; break; }
    const double rho = (cost - cost_c) / model_change;
    if (rho > K.min_rel_decrease) {
      double a2 = 0.0, z1 = 0.0, z2 = 0.0;
      for (int i = tid; i < 2 * C.Nc; i += kCtaThreads) {
        const double v = C.xc[i];
        C.x[i] = v;
        if (C.freeof[i >> 1] >= 0) a2 += v * v;
      }
      block_sum3(C, a2, z1, z2);
      x_norm = sqrt(a2);
      cost = cost_c;
      gmax = cta_assemble<false>(C, false, K);
      success = true;
      const double t = 2.0 * rho - 1.0;
      radius = fmin(K.radius_max, radius / fmax(1.0 / 3.0, 1.0 - t * t * t));
      nu = 2.0;
    } else {
      radius /= nu;
      nu *= 2.0;
    }
  }
  // (not after FAILURE: Ceres only commits a usable solution, solver.cc Minimize / IsSolutionUsable)
  if (term != LFR_TERM_FAILURE) {
    for (int i = tid; i < C.n; i += kCtaThreads) {
      const int l = C.lof[i >> 1];
      P.positions_out[2 * (size_t)C.node[l] + (i & 1)] = C.x[2 * l + (i & 1)];
    }
  }
  if (tid == 0) {
    P.st_iter[c] = iter;
    P.st_term[c] = term;
    P.st_cost0[c] = cost0;
    P.st_cost1[c] = cost;
    P.st_ls[c] = ls_steps;
    if (P.st_cycles) {
      unsigned long long* o = P.st_cycles + 8 * (size_t)c;
      o[0] = o[1] = o[2] = 0;
      o[3] = n_maxit;
      o[5] = (unsigned long long)__double_as_longlong(worst_res);
      o[4] = cg_iters;
      o[1] = 1;  // marks a CTA-tier component
      o[6] = (unsigned long long)ls_steps << 32;
      o[7] = 0;
    }
  }
}

}  // namespace lfr
