// lfr_capi.cu — the C ABI of include/lfr.h for the B200 backend: plan
// construction (problem -> HBM, size-bucketed launch schedule), the solve
// launches, and result download.  No CPU fallback: every entry point that
// computes fails with LFR_ENODEV / LFR_ECUDA when no device is usable.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "lfr_solve_warp.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

#define LFR_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t err__ = (call);                                                          \
    if (err__ != cudaSuccess) {                                                          \
      const int code__ = (err__ == cudaErrorNoDevice || err__ == cudaErrorInsufficientDriver || \
                          err__ == cudaErrorInvalidDevice)                               \
                             ? LFR_ENODEV                                                \
                             : (err__ == cudaErrorMemoryAllocation ? LFR_ENOMEM : LFR_ECUDA); \
      return fail(code__, std::string(#call) + ": " + cudaGetErrorString(err__));        \
    }                                                                                    \
  } while (0)

constexpr int kMaxSmemPerBlock = 227 * 1024;

lfr::DevConsts make_consts(const lfr_options& o) {
  lfr::DevConsts K;
  K.bound = o.bound;
  K.cauchy_b = o.cauchy_a * o.cauchy_a;
  K.cauchy_c = 1.0 / K.cauchy_b;
  K.tukey_a2 = o.tukey_a * o.tukey_a;
  K.tukey_inv_a2 = 1.0 / K.tukey_a2;
  if (o.tukey_variant == 2) {
    K.tukey_rho0 = K.tukey_a2 / 3.0;
    K.tukey_rho1 = 1.0;
  } else {
    K.tukey_rho0 = K.tukey_a2 / 6.0;
    K.tukey_rho1 = 0.5;
  }
  K.f_tol = o.function_tolerance;
  K.g_tol = o.gradient_tolerance;
  K.p_tol = o.parameter_tolerance;
  K.radius0 = o.initial_trust_region_radius;
  K.radius_max = o.max_trust_region_radius;
  K.radius_min = o.min_trust_region_radius;
  K.min_rel_decrease = o.min_relative_decrease;
  K.min_diag = o.min_lm_diagonal;
  K.max_diag = o.max_lm_diagonal;
  K.ls_suff = o.line_search_sufficient_function_decrease;
  K.ls_max_contraction = o.max_line_search_step_contraction;
  K.ls_min_contraction = o.min_line_search_step_contraction;
  K.ls_min_step = o.min_line_search_step_size;
  K.max_iter = o.max_num_iterations;
  K.max_invalid = o.max_num_consecutive_invalid_steps;
  K.max_ls_iter = o.max_num_line_search_step_size_iterations;
  K.linear_solver = o.linear_solver;
  return K;
}

struct Bucket {
  uint32_t* d_list = nullptr;
  uint32_t n = 0;
  int emax = 0, ncmax = 0, n2max = 0, smem_per_warp = 0, warps = 4;
};

int validate(const lfr_problem* p) {
  if (!p) return fail(LFR_EINVAL, "problem is NULL");
  if (p->n_nodes && (!p->row_ptr || !p->track || !p->comp || !p->is_root))
    return fail(LFR_EINVAL, "NULL per-node array");
  if (p->n_components && (!p->comp_ptr || !p->comp_nodes)) return fail(LFR_EINVAL, "NULL component list");
  if (p->n_nodes && p->row_ptr[p->n_nodes] != p->n_edges) return fail(LFR_EINVAL, "row_ptr[n_nodes] != n_edges");
  if (p->n_edges && !p->edges) return fail(LFR_EINVAL, "edges is NULL");
  if (p->n_edges >= (1ull << 32)) return fail(LFR_EUNSUPPORTED, "more than 2^32 directed edges");
  for (uint32_t v = 0; v < p->n_nodes; ++v) {
    if (p->row_ptr[v + 1] < p->row_ptr[v]) return fail(LFR_EINVAL, "row_ptr not monotone");
    for (uint32_t e = p->row_ptr[v]; e < p->row_ptr[v + 1]; ++e) {
      if (p->edges[e].dst >= p->n_nodes) return fail(LFR_EINVAL, "edge dst out of range");
      if (p->edges[e].dst == v) return fail(LFR_EINVAL, "self edge (Ceres rejects duplicate parameter blocks)");
    }
  }
  for (uint32_t c = 0; c < p->n_components; ++c)
    if (p->comp_ptr[c + 1] < p->comp_ptr[c]) return fail(LFR_EINVAL, "comp_ptr not monotone");
  const uint32_t tot = p->n_components ? p->comp_ptr[p->n_components] : 0;
  for (uint32_t i = 0; i < tot; ++i)
    if (p->comp_nodes[i] >= p->n_nodes) return fail(LFR_EINVAL, "comp_nodes out of range");
  return LFR_OK;
}

template <typename T>
int upload(T** d, const T* h, size_t count, cudaStream_t s) {
  *d = nullptr;
  if (count == 0) return LFR_OK;
  LFR_CUDA(cudaMalloc((void**)d, count * sizeof(T)));
  LFR_CUDA(cudaMemcpyAsync(*d, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  return LFR_OK;
}

}  // namespace

struct lfr_plan {
  int device = 0;
  lfr_options opt;
  lfr::DevConsts K;
  uint32_t N = 0, C = 0;
  uint64_t E = 0;
  uint32_t total_slots = 0;
  uint32_t* d_row_ptr = nullptr;
  lfr_edge* d_edges = nullptr;
  uint32_t* d_track = nullptr;
  uint32_t* d_comp = nullptr;
  uint8_t* d_is_root = nullptr;
  uint32_t* d_comp_ptr = nullptr;
  uint32_t* d_comp_nodes = nullptr;
  uint32_t* d_local_of = nullptr;
  double* d_pos = nullptr;
  double* d_pos_init = nullptr;
  int32_t* d_iter = nullptr;
  int32_t* d_term = nullptr;
  double* d_cost0 = nullptr;
  double* d_cost1 = nullptr;
  uint32_t* d_ls = nullptr;
  uint32_t* d_kept = nullptr;
  std::vector<Bucket> buckets;
  std::vector<uint32_t> comp_size;  // nodes per dispatch slot
  uint32_t n_solved = 0;
  bool solved_once = false;

  lfr::DevProblem dev() const {
    lfr::DevProblem P;
    P.row_ptr = d_row_ptr;
    P.edges = reinterpret_cast<const float4*>(d_edges);
    P.track = d_track;
    P.comp = d_comp;
    P.is_root = d_is_root;
    P.comp_ptr = d_comp_ptr;
    P.comp_nodes = d_comp_nodes;
    P.local_of = d_local_of;
    P.positions = d_pos;
    P.st_iter = d_iter;
    P.st_term = d_term;
    P.st_cost0 = d_cost0;
    P.st_cost1 = d_cost1;
    P.st_ls = d_ls;
    P.st_kept = d_kept;
    return P;
  }
};

namespace {

void free_plan(lfr_plan* pl) {
  if (!pl) return;
  cudaFree(pl->d_row_ptr);
  cudaFree(pl->d_edges);
  cudaFree(pl->d_track);
  cudaFree(pl->d_comp);
  cudaFree(pl->d_is_root);
  cudaFree(pl->d_comp_ptr);
  cudaFree(pl->d_comp_nodes);
  cudaFree(pl->d_local_of);
  cudaFree(pl->d_pos);
  cudaFree(pl->d_pos_init);
  cudaFree(pl->d_iter);
  cudaFree(pl->d_term);
  cudaFree(pl->d_cost0);
  cudaFree(pl->d_cost1);
  cudaFree(pl->d_ls);
  cudaFree(pl->d_kept);
  for (Bucket& b : pl->buckets) cudaFree(b.d_list);
  delete pl;
}

// Size-bucketed schedule: components are grouped by the shared memory one warp
// needs for them, so small tracks run at high occupancy and the few large
// components do not dictate the carve-up of everyone else.
int build_buckets(lfr_plan* pl, const lfr_problem* p, cudaStream_t s) {
  static const int kClass[] = {2048, 3072, 4096, 6144, 8192, 12288, 16384, 24576, 32768, 57344, kMaxSmemPerBlock};
  const int n_class = sizeof(kClass) / sizeof(kClass[0]);
  std::vector<std::vector<uint32_t>> members(n_class);
  std::vector<Bucket> caps(n_class);
  pl->comp_size.resize(p->n_components);
  pl->n_solved = 0;
  for (uint32_t c = 0; c < p->n_components; ++c) {
    const uint32_t beg = p->comp_ptr[c], end = p->comp_ptr[c + 1];
    const uint32_t nc = end - beg;
    pl->comp_size[c] = nc;
    if (nc <= 1) continue;  // solve.cc:619-622
    ++pl->n_solved;
    uint64_t eup = 0;
    uint32_t nfree = 0;
    for (uint32_t i = beg; i < end; ++i) {
      const uint32_t v = p->comp_nodes[i];
      eup += p->row_ptr[v + 1] - p->row_ptr[v];
      nfree += p->is_root[v] ? 0 : 1;
    }
    const int n2 = 2 * (int)nfree;
    if (n2 > lfr::kMaxWarpN2 || nc > (uint32_t)lfr::kMaxWarpNodes || eup > 65535) {
      char buf[160];
      snprintf(buf, sizeof buf,
               "component %u (nodes=%u, unknowns=%d, out-edges=%llu) exceeds the warp-tier caps", c, nc, n2,
               (unsigned long long)eup);
      return fail(LFR_EUNSUPPORTED, buf);
    }
    const int e = std::max<int>(1, (int)eup);
    const lfr::WarpLayout L(e, (int)nc, std::max(n2, 2));
    int k = 0;
    while (k < n_class && L.total > kClass[k]) ++k;
    if (k == n_class) return fail(LFR_EUNSUPPORTED, "component needs more shared memory than one SM has");
    members[k].push_back(c);
    caps[k].emax = std::max(caps[k].emax, e);
    caps[k].ncmax = std::max(caps[k].ncmax, (int)nc);
    caps[k].n2max = std::max(caps[k].n2max, std::max(n2, 2));
  }
  for (int k = 0; k < n_class; ++k) {
    if (members[k].empty()) continue;
    Bucket b = caps[k];
    b.n = (uint32_t)members[k].size();
    const lfr::WarpLayout L(b.emax, b.ncmax, b.n2max);
    b.smem_per_warp = L.total;
    if (b.smem_per_warp > kMaxSmemPerBlock) return fail(LFR_EUNSUPPORTED, "bucket exceeds shared memory");
    b.warps = (4 * b.smem_per_warp <= kMaxSmemPerBlock) ? 4 : 1;
    int rc = upload(&b.d_list, members[k].data(), members[k].size(), s);
    if (rc) return rc;
    pl->buckets.push_back(b);
  }
  return LFR_OK;
}

int launch_solve(lfr_plan* pl, cudaStream_t s) {
  LFR_CUDA(cudaMemcpyAsync(pl->d_pos, pl->d_pos_init, sizeof(double) * 2 * (size_t)pl->N,
                           cudaMemcpyDeviceToDevice, s));
  const lfr::DevProblem P = pl->dev();
  for (const Bucket& b : pl->buckets) {
    lfr::WarpBucket wb;
    wb.list = b.d_list;
    wb.n = b.n;
    wb.emax = b.emax;
    wb.ncmax = b.ncmax;
    wb.n2max = b.n2max;
    wb.smem_per_warp = b.smem_per_warp;
    const size_t smem = (size_t)b.smem_per_warp * b.warps;
    const unsigned grid = (b.n + b.warps - 1) / b.warps;
    if (b.warps == 4) {
      LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kMaxSmemPerBlock));
      lfr::solve_warp_kernel<4><<<grid, 128, smem, s>>>(P, pl->K, wb);
    } else {
      LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kMaxSmemPerBlock));
      lfr::solve_warp_kernel<1><<<grid, 32, smem, s>>>(P, pl->K, wb);
    }
    LFR_CUDA(cudaGetLastError());
  }
  pl->solved_once = true;
  return LFR_OK;
}

}  // namespace

extern "C" {

int lfr_abi_version(void) { return LFR_ABI_VERSION; }
const char* lfr_backend(void) { return "b200"; }
const char* lfr_last_error(void) { return g_last_error.c_str(); }

void lfr_options_default(lfr_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof *o);
  o->bound = 1.0;                 // solve.cc:89
  o->cauchy_a = 0.25;             // solve.cc:111
  o->tukey_a = 0.0625;            // solve.cc:120
  o->tukey_variant = 1;
  o->max_num_iterations = 100;    // solve.cc:149
  o->max_num_consecutive_invalid_steps = 10;  // solve.cc:151
  o->max_num_line_search_step_size_iterations = 20;
  o->function_tolerance = 1e-4;   // solve.cc:152
  o->gradient_tolerance = 1e-8;   // solve.cc:153
  o->parameter_tolerance = 1e-4;  // solve.cc:154
  o->initial_trust_region_radius = 1e4;
  o->max_trust_region_radius = 1e16;
  o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3;
  o->min_lm_diagonal = 1e-6;
  o->max_lm_diagonal = 1e32;
  o->line_search_sufficient_function_decrease = 1e-4;
  o->max_line_search_step_contraction = 1e-3;
  o->min_line_search_step_contraction = 0.6;
  o->min_line_search_step_size = 1e-9;
  o->n_threads = 8;
  o->device = 0;
  o->linear_solver = 0;
}

int lfr_plan_create(const lfr_problem* p, const lfr_options* opt, const double* initial_positions,
                    lfr_plan** out) {
  if (!out) return fail(LFR_EINVAL, "out is NULL");
  *out = nullptr;
  int rc = validate(p);
  if (rc) return rc;
  lfr_options o;
  if (opt) o = *opt; else lfr_options_default(&o);
  int n_dev = 0;
  LFR_CUDA(cudaGetDeviceCount(&n_dev));
  if (n_dev == 0) return fail(LFR_ENODEV, "no CUDA device");
  if (o.device < 0 || o.device >= n_dev) return fail(LFR_ENODEV, "device ordinal out of range");
  LFR_CUDA(cudaSetDevice(o.device));
  lfr_plan* pl = new lfr_plan();
  pl->device = o.device;
  pl->opt = o;
  pl->K = make_consts(o);
  pl->N = p->n_nodes;
  pl->C = p->n_components;
  pl->E = p->n_edges;
  pl->total_slots = p->n_components ? p->comp_ptr[p->n_components] : 0;
  cudaStream_t s = 0;
#define LFR_TRY(expr)          \
  do {                         \
    const int rc__ = (expr);   \
    if (rc__) {                \
      free_plan(pl);           \
      return rc__;             \
    }                          \
  } while (0)
  LFR_TRY(upload(&pl->d_row_ptr, p->row_ptr, (size_t)p->n_nodes + 1, s));
  LFR_TRY(upload(&pl->d_edges, p->edges, (size_t)p->n_edges, s));
  LFR_TRY(upload(&pl->d_track, p->track, (size_t)p->n_nodes, s));
  LFR_TRY(upload(&pl->d_comp, p->comp, (size_t)p->n_nodes, s));
  LFR_TRY(upload(&pl->d_is_root, p->is_root, (size_t)p->n_nodes, s));
  LFR_TRY(upload(&pl->d_comp_ptr, p->comp_ptr, (size_t)p->n_components + 1, s));
  LFR_TRY(upload(&pl->d_comp_nodes, p->comp_nodes, (size_t)pl->total_slots, s));
  auto alloc = [&](void** d, size_t bytes) -> int {
    *d = nullptr;
    if (bytes == 0) return LFR_OK;
    LFR_CUDA(cudaMalloc(d, bytes));
    LFR_CUDA(cudaMemsetAsync(*d, 0, bytes, s));
    return LFR_OK;
  };
  LFR_TRY(alloc((void**)&pl->d_local_of, sizeof(uint32_t) * (size_t)pl->N));
  LFR_TRY(alloc((void**)&pl->d_pos, sizeof(double) * 2 * (size_t)pl->N));
  LFR_TRY(alloc((void**)&pl->d_pos_init, sizeof(double) * 2 * (size_t)pl->N));
  LFR_TRY(alloc((void**)&pl->d_iter, sizeof(int32_t) * (size_t)pl->C));
  LFR_TRY(alloc((void**)&pl->d_term, sizeof(int32_t) * (size_t)pl->C));
  LFR_TRY(alloc((void**)&pl->d_cost0, sizeof(double) * (size_t)pl->C));
  LFR_TRY(alloc((void**)&pl->d_cost1, sizeof(double) * (size_t)pl->C));
  LFR_TRY(alloc((void**)&pl->d_ls, sizeof(uint32_t) * (size_t)pl->C));
  LFR_TRY(alloc((void**)&pl->d_kept, sizeof(uint32_t) * (size_t)pl->C));
  if (initial_positions && pl->N) {
    cudaError_t e = cudaMemcpyAsync(pl->d_pos_init, initial_positions, sizeof(double) * 2 * (size_t)pl->N,
                                    cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) {
      free_plan(pl);
      return fail(LFR_ECUDA, cudaGetErrorString(e));
    }
  }
  LFR_TRY(build_buckets(pl, p, s));
  if (pl->total_slots) {
    lfr::local_index_kernel<<<(pl->total_slots + 255) / 256, 256, 0, s>>>(
        pl->d_comp_ptr, pl->d_comp_nodes, pl->C, pl->total_slots, pl->d_local_of);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      free_plan(pl);
      return fail(LFR_ECUDA, cudaGetErrorString(e));
    }
  }
  {
    cudaError_t e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      free_plan(pl);
      return fail(LFR_ECUDA, cudaGetErrorString(e));
    }
  }
#undef LFR_TRY
  *out = pl;
  return LFR_OK;
}

int lfr_plan_solve(lfr_plan* pl, void* stream) {
  if (!pl) return fail(LFR_EINVAL, "plan is NULL");
  LFR_CUDA(cudaSetDevice(pl->device));
  return launch_solve(pl, (cudaStream_t)stream);
}

int lfr_plan_num_launches(const lfr_plan* pl) {
  // kernels only: the device-to-device reset of `positions` is a copy, not a kernel
  return pl ? (int)pl->buckets.size() : 0;
}

int lfr_plan_download(lfr_plan* pl, void* stream, double* positions, lfr_stats* st) {
  if (!pl) return fail(LFR_EINVAL, "plan is NULL");
  LFR_CUDA(cudaSetDevice(pl->device));
  cudaStream_t s = (cudaStream_t)stream;
  if (positions && pl->N)
    LFR_CUDA(cudaMemcpyAsync(positions, pl->d_pos, sizeof(double) * 2 * (size_t)pl->N, cudaMemcpyDeviceToHost, s));
  std::vector<int32_t> it;
  std::vector<uint32_t> ls;
  if (st && pl->C) {
    it.resize(pl->C);
    ls.resize(pl->C);
    LFR_CUDA(cudaMemcpyAsync(it.data(), pl->d_iter, sizeof(int32_t) * pl->C, cudaMemcpyDeviceToHost, s));
    LFR_CUDA(cudaMemcpyAsync(ls.data(), pl->d_ls, sizeof(uint32_t) * pl->C, cudaMemcpyDeviceToHost, s));
    if (st->termination)
      LFR_CUDA(cudaMemcpyAsync(st->termination, pl->d_term, sizeof(int32_t) * pl->C, cudaMemcpyDeviceToHost, s));
    if (st->initial_cost)
      LFR_CUDA(cudaMemcpyAsync(st->initial_cost, pl->d_cost0, sizeof(double) * pl->C, cudaMemcpyDeviceToHost, s));
    if (st->final_cost)
      LFR_CUDA(cudaMemcpyAsync(st->final_cost, pl->d_cost1, sizeof(double) * pl->C, cudaMemcpyDeviceToHost, s));
  }
  LFR_CUDA(cudaStreamSynchronize(s));
  if (st) {
    uint64_t ti = 0, tl = 0;
    for (uint32_t c = 0; c < pl->C; ++c) {
      ti += (uint64_t)it[c];
      tl += ls[c];
    }
    if (st->iterations && pl->C) std::memcpy(st->iterations, it.data(), sizeof(int32_t) * pl->C);
    st->total_iterations = ti;
    st->total_line_search_steps = tl;
    st->n_solved = pl->n_solved;
    st->n_kernel_launches = (uint32_t)pl->buckets.size();
  }
  return LFR_OK;
}

int lfr_plan_traffic(lfr_plan* pl, void* stream, uint64_t* algorithmic_bytes, uint64_t* one_pass_bytes) {
  if (!pl) return fail(LFR_EINVAL, "plan is NULL");
  LFR_CUDA(cudaSetDevice(pl->device));
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<int32_t> it(pl->C);
  std::vector<uint32_t> kept(pl->C);
  if (pl->C) {
    LFR_CUDA(cudaMemcpyAsync(it.data(), pl->d_iter, sizeof(int32_t) * pl->C, cudaMemcpyDeviceToHost, s));
    LFR_CUDA(cudaMemcpyAsync(kept.data(), pl->d_kept, sizeof(uint32_t) * pl->C, cudaMemcpyDeviceToHost, s));
  }
  LFR_CUDA(cudaStreamSynchronize(s));
  uint64_t alg = 0, one = 0;
  for (uint32_t c = 0; c < pl->C; ++c) {
    if (pl->comp_size[c] <= 1) continue;
    const uint64_t pass = 80ull * kept[c] + 36ull * pl->comp_size[c];
    one += pass;
    alg += pass * (uint64_t)std::max(it[c], 1);
  }
  if (algorithmic_bytes) *algorithmic_bytes = alg;
  if (one_pass_bytes) *one_pass_bytes = one;
  return LFR_OK;
}

void lfr_plan_destroy(lfr_plan* pl) {
  if (!pl) return;
  cudaSetDevice(pl->device);
  free_plan(pl);
}

int lfr_solve(const lfr_problem* p, const lfr_options* opt, double* positions, lfr_stats* st) {
  if (p && p->n_nodes && !positions) return fail(LFR_EINVAL, "positions is NULL");
  lfr_plan* pl = nullptr;
  cudaEvent_t ev[4];
  int n_dev = 0;
  LFR_CUDA(cudaGetDeviceCount(&n_dev));
  if (n_dev == 0) return fail(LFR_ENODEV, "no CUDA device");
  LFR_CUDA(cudaSetDevice(opt ? opt->device : 0));
  for (int i = 0; i < 4; ++i) LFR_CUDA(cudaEventCreate(&ev[i]));
  cudaStream_t s = 0;
  LFR_CUDA(cudaEventRecord(ev[0], s));
  int rc = lfr_plan_create(p, opt, positions, &pl);
  if (rc) return rc;
  cudaEventRecord(ev[1], s);
  rc = launch_solve(pl, s);
  if (rc) {
    free_plan(pl);
    return rc;
  }
  cudaEventRecord(ev[2], s);
  rc = lfr_plan_download(pl, s, positions, st);
  cudaEventRecord(ev[3], s);
  cudaEventSynchronize(ev[3]);
  if (st && rc == LFR_OK) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ev[0], ev[1]);
    cudaEventElapsedTime(&b, ev[1], ev[2]);
    cudaEventElapsedTime(&c, ev[2], ev[3]);
    st->h2d_ms = a;
    st->kernel_ms = b;
    st->d2h_ms = c;
    st->total_ms = (double)a + b + c;
  }
  for (int i = 0; i < 4; ++i) cudaEventDestroy(ev[i]);
  free_plan(pl);
  return rc;
}

int lfr_debug_edge_eval(const lfr_edge* edges, const uint8_t* kind, uint64_t n, const double* xs,
                        const double* xd, const lfr_options* opt, double* r, double* jac, double* rho) {
  lfr_options o;
  if (opt) o = *opt; else lfr_options_default(&o);
  int n_dev = 0;
  LFR_CUDA(cudaGetDeviceCount(&n_dev));
  if (n_dev == 0) return fail(LFR_ENODEV, "no CUDA device");
  LFR_CUDA(cudaSetDevice(o.device));
  if (n == 0) return LFR_OK;
  lfr_edge* d_e = nullptr;
  uint8_t* d_k = nullptr;
  double *d_xs = nullptr, *d_xd = nullptr, *d_r = nullptr, *d_j = nullptr, *d_rho = nullptr;
  LFR_CUDA(cudaMalloc((void**)&d_e, n * sizeof(lfr_edge)));
  LFR_CUDA(cudaMalloc((void**)&d_k, n));
  LFR_CUDA(cudaMalloc((void**)&d_xs, n * 16));
  LFR_CUDA(cudaMalloc((void**)&d_xd, n * 16));
  LFR_CUDA(cudaMalloc((void**)&d_r, n * 16));
  LFR_CUDA(cudaMalloc((void**)&d_j, n * 32));
  LFR_CUDA(cudaMalloc((void**)&d_rho, n * 24));
  LFR_CUDA(cudaMemcpy(d_e, edges, n * sizeof(lfr_edge), cudaMemcpyHostToDevice));
  LFR_CUDA(cudaMemcpy(d_k, kind, n, cudaMemcpyHostToDevice));
  LFR_CUDA(cudaMemcpy(d_xs, xs, n * 16, cudaMemcpyHostToDevice));
  LFR_CUDA(cudaMemcpy(d_xd, xd, n * 16, cudaMemcpyHostToDevice));
  lfr::edge_eval_kernel<<<(unsigned)((n + 127) / 128), 128>>>(reinterpret_cast<const float4*>(d_e), d_k, n, d_xs,
                                                             d_xd, make_consts(o), d_r, d_j, d_rho);
  LFR_CUDA(cudaGetLastError());
  LFR_CUDA(cudaMemcpy(r, d_r, n * 16, cudaMemcpyDeviceToHost));
  LFR_CUDA(cudaMemcpy(jac, d_j, n * 32, cudaMemcpyDeviceToHost));
  LFR_CUDA(cudaMemcpy(rho, d_rho, n * 24, cudaMemcpyDeviceToHost));
  cudaFree(d_e); cudaFree(d_k); cudaFree(d_xs); cudaFree(d_xd); cudaFree(d_r); cudaFree(d_j); cudaFree(d_rho);
  return LFR_OK;
}

}  // extern "C"
