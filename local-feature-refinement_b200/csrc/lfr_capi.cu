// lfr_capi.cu — the C ABI of include/lfr.h for the B200 backend: plan
// construction (problem -> HBM, size-bucketed launch schedule), the solve
// launches, and result download.  No CPU fallback: every entry point that
// computes fails with LFR_ENODEV / LFR_ECUDA when no device is usable.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "lfr_solve_cta.cuh"
#include "lfr_solve_tile.cuh"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}

int cuda_code(cudaError_t e) {
  if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInvalidDevice) return LFR_ENODEV;
  if (e == cudaErrorMemoryAllocation) return LFR_ENOMEM;
  return LFR_ECUDA;
}

#define LFR_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t err__ = (call);                                                     \
    if (err__ != cudaSuccess)                                                       \
      return fail(cuda_code(err__), std::string(#call) + ": " + cudaGetErrorString(err__)); \
  } while (0)

#define LFR_TRY(expr)        \
  do {                       \
    const int rc__ = (expr); \
    if (rc__) return rc__;   \
  } while (0)

constexpr int kMaxSmemPerBlock = 227 * 1024;
// Components with more unknowns than this (and <= 32) take the two-warp tile kernel <64, 32> instead of
// the one-warp register kernel; 32 = never.  Overridden per plan by LFR_TILE_FROM_N.
constexpr int kTileFromDefault = 32;
constexpr int kMaxStreams = 12;

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

// Grow-only device buffer: lfr_solve() re-uses its workspace across calls, so a
// steady stream of solves performs no cudaMalloc/cudaFree.
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int reserve(size_t bytes) {
    if (bytes <= cap) return LFR_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) return fail(cuda_code(e), std::string("cudaMalloc: ") + cudaGetErrorString(e));
    cap = want;
    return LFR_OK;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  template <typename T>
  T* as() const { return static_cast<T*>(p); }
};

struct Bucket {
  uint32_t offset = 0;  // into the bucket-list buffer
  uint32_t n = 0;
  int emax = 0, ncmax = 0, n2max = 0, smem_per_warp = 0, warps = 4;
  int variant = 0;  // 0: shared-memory Cholesky kernel (n <= 64); 16 / 32: register Gauss-Jordan kernel (n <= variant)
  bool stages_edges() const { return variant != 0; }  // warp2 / tile tiers pull their edge records into shared memory
};

int validate(const lfr_problem* p) {
  if (!p) return fail(LFR_EINVAL, "problem is NULL");
  if (p->n_nodes && (!p->row_ptr || !p->track || !p->comp || !p->is_root))
    return fail(LFR_EINVAL, "NULL per-node array");
  if (p->n_components && (!p->comp_ptr || !p->comp_nodes)) return fail(LFR_EINVAL, "NULL component list");
  if (p->n_nodes && p->row_ptr[p->n_nodes] != p->n_edges) return fail(LFR_EINVAL, "row_ptr[n_nodes] != n_edges");
  if (p->n_edges && !p->edges) return fail(LFR_EINVAL, "edges is NULL");
  if (p->n_edges >= (1ull << 32)) return fail(LFR_EUNSUPPORTED, "more than 2^32 directed edges");
  // per-edge checks (dst range, self edges) run on the device while the edges are
  // staged (solve_warp_kernel), so the host never walks the 80-byte records.
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
  DevBuf row_ptr, edges, track, comp, is_root, comp_ptr, comp_nodes, local_of, pos, pos_init, stats, cycles, times, lists;
  // `stats` is one block (one memset, one D2H copy): cost0[Cp] cost1[Cp] iter[Cp] term[Cp] ls[Cp] kept[Cp] err[2]
  uint32_t Cp = 0;               // C rounded up to an even count
  bool pos_is_staged = false;    // lfr_solve(): the start point was uploaded straight into `pos`
  void* h_stage = nullptr;       // pinned host staging for the stats block
  size_t h_stage_cap = 0;
  double* d_cost0() const { return stats.as<double>(); }
  double* d_cost1() const { return stats.as<double>() + Cp; }
  int32_t* d_iter() const { return reinterpret_cast<int32_t*>(stats.as<double>() + 2 * (size_t)Cp); }
  int32_t* d_term() const { return d_iter() + Cp; }
  uint32_t* d_ls() const { return reinterpret_cast<uint32_t*>(d_term() + Cp); }
  uint32_t* d_kept() const { return d_ls() + Cp; }
  int* d_err() const { return reinterpret_cast<int*>(d_kept() + Cp); }
  size_t stats_bytes() const { return 16 * (size_t)Cp + 16 * (size_t)Cp + 8; }
  std::vector<Bucket> buckets;
  std::vector<uint32_t> comp_size;  // nodes per dispatch slot
  std::vector<uint32_t> list_host;
  // CTA tier (block-Jacobi PCG): components with more than kMaxWarpN2 unknowns, or all with linear_solver = 2
  std::vector<uint32_t> large_slots;
  uint32_t n_large = 0;
  DevBuf L_comps, L_eidx, L_meta, L_inlist, L_twin, L_fdst, L_bmat, L_scr, L_q, L_node, L_outptr, L_inptr, L_freeof, L_x, L_xc, L_lof, L_vec;
  uint64_t L_total_free = 0;
  uint32_t L_max_free = 0;
  uint32_t n_solved = 0;
  bool profile = false;
  // lfr_solve() zero-copy: staging tiers read the caller's pinned edge array / write the caller's
  // pinned positions directly (device-side addresses of those host buffers), nullptr = through HBM
  const float4* zc_edges = nullptr;
  double* zc_positions = nullptr;
  const uint8_t* slot_owner = nullptr;  // lfr_solve_multi(): owner[c] of every dispatch slot, this plan solves owner == owner_id
  uint8_t owner_id = 0;
  bool edges_in_hbm = false;       // the edge array was (or is being) copied to `edges`
  bool needs_hbm_edges = false;    // some bucket (smem-Cholesky warp tier, CTA tier) reads edges from global memory
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_edges = nullptr;  // bulk edge copy done (only the non-staging tiers wait for it)
  cudaEvent_t ev_small = nullptr;  // fork / join of the second upload stream
  cudaStream_t streams[kMaxStreams] = {};
  cudaEvent_t ev_fork = nullptr, ev_join[kMaxStreams] = {};
  int n_streams = 0;

  lfr::DevProblem dev() const {
    lfr::DevProblem P;
    P.n_nodes = N;
    P.row_ptr = row_ptr.as<uint32_t>();
    P.edges = edges.as<float4>();
    P.track = track.as<uint32_t>();
    P.comp = comp.as<uint32_t>();
    P.is_root = is_root.as<uint8_t>();
    P.comp_ptr = comp_ptr.as<uint32_t>();
    P.comp_nodes = comp_nodes.as<uint32_t>();
    P.local_of = local_of.as<uint32_t>();
    P.positions = pos.as<double>();
    P.positions_out = zc_positions ? zc_positions : pos.as<double>();
    P.stage_mode = (opt.debug_flags & LFR_DBG_STAGE_LDG) ? 0 : 1;
    P.st_iter = d_iter();
    P.st_term = d_term();
    P.st_cost0 = d_cost0();
    P.st_cost1 = d_cost1();
    P.st_ls = d_ls();
    P.st_kept = d_kept();
    P.st_cycles = profile ? cycles.as<unsigned long long>() : nullptr;
    P.st_times = profile ? times.as<unsigned long long>() : nullptr;
    P.err_flag = d_err();
    return P;
  }
};

namespace {

void free_plan(lfr_plan* pl) {
  if (!pl) return;
  DevBuf* bufs[] = {&pl->row_ptr, &pl->edges, &pl->track, &pl->comp, &pl->is_root, &pl->comp_ptr, &pl->comp_nodes,
                    &pl->local_of, &pl->pos, &pl->pos_init, &pl->stats, &pl->cycles, &pl->times, &pl->lists, &pl->L_comps, &pl->L_eidx, &pl->L_meta,
                    &pl->L_inlist, &pl->L_twin, &pl->L_fdst, &pl->L_bmat, &pl->L_scr, &pl->L_q, &pl->L_node, &pl->L_outptr, &pl->L_inptr, &pl->L_freeof,
                    &pl->L_x, &pl->L_xc, &pl->L_lof, &pl->L_vec};
  for (DevBuf* b : bufs) b->release();
  if (pl->h_stage) cudaFreeHost(pl->h_stage);
  for (int i = 0; i < pl->n_streams; ++i) {
    if (pl->streams[i]) cudaStreamDestroy(pl->streams[i]);
    if (pl->ev_join[i]) cudaEventDestroy(pl->ev_join[i]);
  }
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_edges) cudaEventDestroy(pl->ev_edges);
  if (pl->ev_small) cudaEventDestroy(pl->ev_small);
  if (pl->copy_stream) cudaStreamDestroy(pl->copy_stream);
  delete pl;
}

template <typename T>
int upload(DevBuf* d, const T* h, size_t count, cudaStream_t s) {
  LFR_TRY(d->reserve(std::max<size_t>(count, 1) * sizeof(T)));
  if (count) LFR_CUDA(cudaMemcpyAsync(d->p, h, count * sizeof(T), cudaMemcpyHostToDevice, s));
  return LFR_OK;
}

// Size-bucketed schedule: components are grouped by the shared memory one warp
// needs for them, so small tracks run at high occupancy and the few large
// components do not dictate the carve-up of everyone else.  Buckets are
// launched on concurrent streams, largest components first.
int build_buckets(lfr_plan* pl, const lfr_problem* p) {
  static const int kClass[] = {2048, 3072, 4096, 6144, 8192, 12288, 16384, 24576, 32768, 57344, kMaxSmemPerBlock};
  const int n_class = sizeof(kClass) / sizeof(kClass[0]);
  // register warp kernel (8..32), register tile kernel (132 = 64 threads x NREG 32; 48, 64: 64 threads;
  // 80: 128 threads), smem-Cholesky warp kernel (0)
  static const int kVariant[9] = {8, 16, 24, 32, 132, 48, 64, 80, 0};
  const int kNV = 9, kV1 = 8;
  auto is_tile = [](int vi) { return vi >= 4 && vi <= 7; };
  const int dbg = pl->opt.debug_flags;
  const bool no_tile = (dbg & LFR_DBG_NO_TILE) != 0;
  // components with more than tile_from unknowns (and <= 32) go to the two-warp tile kernel
  // instead of the one-warp kernel (their edge evaluation is split over 64 threads)
  const int tf = (dbg >> LFR_DBG_TILE_FROM_SHIFT) & 0xff;
  const int tile_from = no_tile ? 32 : (tf ? tf : kTileFromDefault);
  std::vector<std::vector<uint32_t>> members(kNV * n_class);
  std::vector<Bucket> caps(kNV * n_class);
  pl->comp_size.resize(p->n_components);
  pl->n_solved = 0;
  pl->buckets.clear();
  pl->list_host.clear();
  const bool force_v1 = (dbg & LFR_DBG_FORCE_SMEM_CHOLESKY) != 0;
  const bool force_pcg = pl->opt.linear_solver == 2;
  pl->large_slots.clear();
  auto layout_bytes = [&](int vi, int e, int nc, int n2) {
    return (vi == kV1) ? lfr::WarpLayout(e, nc, n2).total
                       : (is_tile(vi) ? lfr::TileLayout(e, nc, n2).total : lfr::Warp2Layout(e, nc, n2).total);
  };
  struct Dim { int e, nc, n2; };
  std::vector<Dim> dim(p->n_components, Dim{0, 0, 0});
  for (uint32_t c = 0; c < p->n_components; ++c) {
    const uint32_t beg = p->comp_ptr[c], end = p->comp_ptr[c + 1];
    if (end < beg || end > pl->total_slots) return fail(LFR_EINVAL, "comp_ptr not monotone");
    const uint32_t nc = end - beg;
    pl->comp_size[c] = nc;
    if (nc <= 1) continue;  // solve.cc:619-622
    if (pl->slot_owner && pl->slot_owner[c] != pl->owner_id) continue;  // another device's component
    ++pl->n_solved;
    uint64_t eup = 0;
    uint32_t nfree = 0;
    for (uint32_t i = beg; i < end; ++i) {
      const uint32_t v = p->comp_nodes[i];
      if (v >= p->n_nodes) return fail(LFR_EINVAL, "comp_nodes out of range");
      if (p->row_ptr[v + 1] < p->row_ptr[v]) return fail(LFR_EINVAL, "row_ptr not monotone");
      eup += p->row_ptr[v + 1] - p->row_ptr[v];
      nfree += p->is_root[v] ? 0 : 1;
    }
    const int n2 = std::max(2 * (int)nfree, 2);
    auto to_cta_tier = [&]() -> int {
      if (nc > 16383) return fail(LFR_EUNSUPPORTED, "component with more than 16383 nodes");
      pl->large_slots.push_back(c);
      return LFR_OK;
    };
    if (force_pcg || n2 > lfr::kMaxWarpN2 || nc > (uint32_t)lfr::kMaxWarpNodes || eup > 65535) {
      LFR_TRY(to_cta_tier());
      continue;
    }
    const int e = std::max<int>(1, (int)eup);
    int vi = (n2 <= 8) ? 0 : (n2 <= 16 ? 1 : (n2 <= 24 ? 2 : (n2 <= 32 ? 3 : (n2 <= 48 ? 5 : (n2 <= 64 ? 6 : (n2 <= 80 ? 7 : kV1))))));
    if (vi <= 3 && n2 > tile_from && lfr::TileLayout(e, (int)nc, n2).total <= kMaxSmemPerBlock) vi = 4;
    if (vi >= 5 && is_tile(vi) && (no_tile || lfr::TileLayout(e, (int)nc, n2).total > kMaxSmemPerBlock)) vi = kV1;
    if (force_v1) vi = kV1;
    int need = layout_bytes(vi, e, (int)nc, n2);
    if (need > kMaxSmemPerBlock && vi != kV1) {  // the staged records do not fit: the Cholesky warp kernel reads them from global memory
      vi = kV1;
      need = layout_bytes(vi, e, (int)nc, n2);
    }
    if (need > kMaxSmemPerBlock) {
      // a dense, high-degree component (its per-edge scratch alone exceeds one SM's shared
      // memory): the CTA tier keeps its per-edge data in HBM
      LFR_TRY(to_cta_tier());
      continue;
    }
    int k = 0;
    while (k < n_class && need > kClass[k]) ++k;
    Bucket& cb = caps[vi * n_class + k];
    members[vi * n_class + k].push_back(c);
    dim[c] = Dim{e, (int)nc, n2};
    cb.emax = std::max(cb.emax, e);
    cb.ncmax = std::max(cb.ncmax, (int)nc);
    cb.n2max = std::max(cb.n2max, n2);
  }
  pl->needs_hbm_edges = !pl->large_slots.empty();
  auto emit = [&](int vi, const Bucket& cap, const uint32_t* mem, size_t n_mem) {
    Bucket b = cap;
    b.variant = kVariant[vi];
    b.n = (uint32_t)n_mem;
    b.offset = (uint32_t)pl->list_host.size();
    b.smem_per_warp = layout_bytes(vi, b.emax, b.ncmax, b.n2max);
    b.warps = (4 * b.smem_per_warp <= kMaxSmemPerBlock) ? 4 : 1;
    if (b.variant >= 48) b.warps = 1;  // tile kernels: one component per CTA
    if (b.warps == 1 && b.variant != 0 && b.variant < 48) b.variant = 32;  // only <1,32> is instantiated for single-warp CTAs
    pl->list_host.insert(pl->list_host.end(), mem, mem + n_mem);
    pl->buckets.push_back(b);
    if (!b.stages_edges()) pl->needs_hbm_edges = true;
  };
  for (int k = n_class - 1; k >= 0; --k) {  // largest first
    for (int vi = kNV - 1; vi >= 0; --vi) {
      const std::vector<uint32_t>& mem = members[vi * n_class + k];
      if (mem.empty()) continue;
      const Bucket& cap = caps[vi * n_class + k];
      if (layout_bytes(vi, cap.emax, cap.ncmax, cap.n2max) <= kMaxSmemPerBlock) {
        emit(vi, cap, mem.data(), mem.size());
      } else {
        // every member fits on its own but the union of their maxima does not: one launch each
        for (uint32_t c : mem) {
          Bucket one;
          one.emax = dim[c].e;
          one.ncmax = dim[c].nc;
          one.n2max = dim[c].n2;
          emit(vi, one, &c, 1);
        }
      }
    }
  }
  return LFR_OK;
}

// Host-side preparation of the CTA-tier components (solve.cc:98-143 for each):
// kept-edge lists with local indices and loss kinds, in-edge lists, free-variable
// numbering.  These are the few components too large for one warp.
int prepare_large(lfr_plan* pl, const lfr_problem* p, cudaStream_t s) {
  pl->n_large = (uint32_t)pl->large_slots.size();
  pl->L_total_free = 0;
  pl->L_max_free = 0;
  if (pl->n_large == 0) return LFR_OK;
  std::vector<lfr::CtaComp> comps(pl->n_large);
  std::vector<uint32_t> eidx, meta, inlist, twin, node, outptr, inptr, lof;
  std::vector<int32_t> freeof, fdst;
  std::vector<int32_t> local(p->n_nodes, -1);
  uint64_t e_off = 0, n_off = 0, f_off = 0;
  for (uint32_t k = 0; k < pl->n_large; ++k) {
    const uint32_t c = pl->large_slots[k];
    const uint32_t beg = p->comp_ptr[c], nc = p->comp_ptr[c + 1] - beg;
    for (uint32_t l = 0; l < nc; ++l) local[p->comp_nodes[beg + l]] = (int32_t)l;
    const size_t e0 = eidx.size();
    std::vector<uint32_t> cin(nc, 0), cout(nc, 0);
    outptr.push_back(0);
    for (uint32_t l = 0; l < nc; ++l) {
      const uint32_t v = p->comp_nodes[beg + l];
      node.push_back(v);
      for (uint32_t e = p->row_ptr[v]; e < p->row_ptr[v + 1]; ++e) {
        const uint32_t dst = p->edges[e].dst;
        if (dst >= p->n_nodes || dst == v) return fail(LFR_EINVAL, "edge with dst out of range or a self edge");
        uint32_t kind;
        if (p->track[v] == p->track[dst]) kind = LFR_EDGE_CAUCHY;        // solve.cc:105
        else if (p->comp[v] == p->comp[dst]) kind = LFR_EDGE_TUKEY;      // solve.cc:114
        else continue;                                                   // solve.cc:123
        if (p->is_root[v] && p->is_root[dst]) continue;                  // all-constant block (A.1)
        const int32_t dl = local[dst];
        if (dl < 0) return fail(LFR_EINVAL, "component_idx and nodes_in_component disagree");
        eidx.push_back(e);
        meta.push_back(l | ((uint32_t)dl << 14) | (kind << 28));
        ++cout[l];
        ++cin[dl];
      }
      outptr.push_back((uint32_t)(eidx.size() - e0));
    }
    const uint32_t ec = (uint32_t)(eidx.size() - e0);
    // in-edge lists: counting sort by destination (stable => ascending edge index)
    std::vector<uint32_t> ip(nc + 1, 0);
    for (uint32_t l = 0; l < nc; ++l) ip[l + 1] = ip[l] + cin[l];
    std::vector<uint32_t> fill(ip.begin(), ip.end() - 1);
    inlist.resize(e0 + ec);
    for (uint32_t j = 0; j < ec; ++j) {
      const uint32_t dl = (meta[e0 + j] >> 14) & 0x3fff;
      inlist[e0 + fill[dl]++] = j;
    }
    inptr.insert(inptr.end(), ip.begin(), ip.end());
    uint32_t nf = 0;
    for (uint32_t l = 0; l < nc; ++l) {
      const bool is_free = (cout[l] + cin[l] > 0) && !p->is_root[p->comp_nodes[beg + l]];
      freeof.push_back(is_free ? (int32_t)nf : -1);
      if (is_free) {
        lof.push_back(l);
        ++nf;
      }
    }
    for (uint32_t l = 0; l < nc; ++l) local[p->comp_nodes[beg + l]] = -1;
    // twin (reverse edge) of every kept edge and the destination's free index
    bool regular = true;
    twin.resize(e0 + ec);
    fdst.resize(e0 + ec);
    const size_t op0 = outptr.size() - (nc + 1), fo0 = freeof.size() - nc;
    for (uint32_t j = 0; j < ec; ++j) {
      const uint32_t sl = meta[e0 + j] & 0x3fff, dl = (meta[e0 + j] >> 14) & 0x3fff;
      uint32_t found = 0, tw = j;
      for (uint32_t t = outptr[op0 + dl]; t < outptr[op0 + dl + 1]; ++t)
        if (((meta[e0 + t] >> 14) & 0x3fff) == sl) {
          tw = t;
          ++found;
        }
      if (found != 1) regular = false;
      twin[e0 + j] = tw;
      fdst[e0 + j] = freeof[fo0 + dl];
    }
    lfr::CtaComp& cc = comps[k];
    cc.slot = c;
    cc.Nc = nc;
    cc.Ec = ec;
    cc.nf = nf;
    cc.regular = regular ? 1u : 0u;
    cc.e_off = e_off;
    cc.n_off = n_off;
    cc.f_off = f_off;
    cc.comp_index = k;
    pl->L_max_free = std::max(pl->L_max_free, nf);
    e_off += ec;
    n_off += nc;
    f_off += nf;
  }
  pl->L_total_free = f_off;
  LFR_TRY(upload(&pl->L_comps, comps.data(), comps.size(), s));
  LFR_TRY(upload(&pl->L_eidx, eidx.data(), eidx.size(), s));
  LFR_TRY(upload(&pl->L_meta, meta.data(), meta.size(), s));
  LFR_TRY(upload(&pl->L_inlist, inlist.data(), inlist.size(), s));
  LFR_TRY(upload(&pl->L_twin, twin.data(), twin.size(), s));
  LFR_TRY(upload(&pl->L_fdst, fdst.data(), fdst.size(), s));
  LFR_TRY(pl->L_bmat.reserve(sizeof(double) * 4 * std::max<uint64_t>(e_off, 1)));
  LFR_TRY(upload(&pl->L_node, node.data(), node.size(), s));
  LFR_TRY(upload(&pl->L_outptr, outptr.data(), outptr.size(), s));
  LFR_TRY(upload(&pl->L_inptr, inptr.data(), inptr.size(), s));
  LFR_TRY(upload(&pl->L_freeof, freeof.data(), freeof.size(), s));
  LFR_TRY(upload(&pl->L_lof, lof.data(), lof.size(), s));
  LFR_TRY(pl->L_scr.reserve(sizeof(double) * 7 * std::max<uint64_t>(e_off, 1)));
  LFR_TRY(pl->L_q.reserve(sizeof(double) * 2 * std::max<uint64_t>(e_off, 1)));
  LFR_TRY(pl->L_x.reserve(sizeof(double) * 2 * std::max<uint64_t>(n_off, 1)));
  LFR_TRY(pl->L_xc.reserve(sizeof(double) * 2 * std::max<uint64_t>(n_off, 1)));
  LFR_TRY(pl->L_vec.reserve(sizeof(double) * (2 * lfr::V_COUNT + 9) * std::max<uint64_t>(f_off, 1)));
  // the upload sources above are locals: make sure the copies are done before they go away
  LFR_CUDA(cudaStreamSynchronize(s));
  return LFR_OK;
}

// Device-side address of a page-locked (cudaHostAlloc / cudaHostRegister) host buffer, or nullptr
// when the memory is pageable (or not 16-byte aligned, which the 128-bit accesses need).
void* device_view_of_pinned(const void* host_ptr) {
  if (!host_ptr || (reinterpret_cast<uintptr_t>(host_ptr) & 15u)) return nullptr;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host_ptr) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return a.devicePointer;
}

// (Re)fill a plan from host arrays: H2D copies + schedule.  Buffers only grow.
// zc_edges / zc_positions: device-side views of the caller's pinned buffers (lfr_solve() only):
// the staging tiers then read the edge records / write the results there, and the 80-byte records
// — 95 % of the input bytes — are copied to HBM only if some non-staging tier needs them.
int fill_plan(lfr_plan* pl, const lfr_problem* p, const lfr_options& o, const double* initial_positions,
              cudaStream_t s, bool stage_positions_directly = false, const float4* zc_edges = nullptr,
              double* zc_positions = nullptr) {
  pl->opt = o;
  pl->K = make_consts(o);
  pl->N = p->n_nodes;
  pl->C = p->n_components;
  pl->E = p->n_edges;
  pl->total_slots = p->n_components ? p->comp_ptr[p->n_components] : 0;
  pl->profile = (o.debug_flags & LFR_DBG_PROFILE) != 0;
  pl->zc_edges = zc_edges;
  pl->zc_positions = zc_positions;
  pl->edges_in_hbm = false;
  if (!zc_edges) {
    LFR_TRY(upload(&pl->edges, p->edges, (size_t)p->n_edges, s));  // the bulk first: the DMA runs while the host schedules
    pl->edges_in_hbm = true;
  }
  // the small per-node arrays gate the first launch: with zero-copy edges they ARE the upload, so they
  // go out on two streams (two copy engines) instead of queueing behind each other
  cudaStream_t s2 = s;
  if (zc_edges) {
    if (!pl->copy_stream) LFR_CUDA(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
    if (!pl->ev_small) LFR_CUDA(cudaEventCreateWithFlags(&pl->ev_small, cudaEventDisableTiming));
    s2 = pl->copy_stream;
    LFR_CUDA(cudaEventRecord(pl->ev_small, s));          // orders s2's copies after whatever `s` ran before
    LFR_CUDA(cudaStreamWaitEvent(s2, pl->ev_small, 0));
  }
  LFR_TRY(upload(&pl->row_ptr, p->row_ptr, (size_t)p->n_nodes + 1, s));
  LFR_TRY(upload(&pl->track, p->track, (size_t)p->n_nodes, s2));
  LFR_TRY(upload(&pl->comp, p->comp, (size_t)p->n_nodes, s2));
  LFR_TRY(upload(&pl->is_root, p->is_root, (size_t)p->n_nodes, s2));
  LFR_TRY(upload(&pl->comp_ptr, p->comp_ptr, (size_t)p->n_components + 1, s));
  LFR_TRY(upload(&pl->comp_nodes, p->comp_nodes, (size_t)pl->total_slots, s));
  if (s2 != s) {
    LFR_CUDA(cudaEventRecord(pl->ev_small, s2));
    LFR_CUDA(cudaStreamWaitEvent(s, pl->ev_small, 0));
  }
  const size_t N = std::max<size_t>(pl->N, 1), C = std::max<size_t>(pl->C, 1);
  LFR_TRY(pl->local_of.reserve(sizeof(uint32_t) * N));
  LFR_TRY(pl->pos.reserve(sizeof(double) * 2 * N));
  pl->Cp = (uint32_t)((C + 1) & ~(size_t)1);
  LFR_TRY(pl->stats.reserve(pl->stats_bytes()));
  if (pl->profile) {
    LFR_TRY(pl->cycles.reserve(sizeof(unsigned long long) * 8 * C));
    LFR_TRY(pl->times.reserve(sizeof(unsigned long long) * 2 * C));
    LFR_CUDA(cudaMemsetAsync(pl->cycles.p, 0, sizeof(unsigned long long) * 8 * C, s));
    LFR_CUDA(cudaMemsetAsync(pl->times.p, 0, sizeof(unsigned long long) * 2 * C, s));
  }
  // per-slot stats default to "skipped" (size-1 components never run); also clears the error flag
  LFR_CUDA(cudaMemsetAsync(pl->stats.p, 0, pl->stats_bytes(), s));
  if (stage_positions_directly) {
    // lfr_solve(): the caller's start point goes straight into the working array
    if (pl->N)
      LFR_CUDA(cudaMemcpyAsync(pl->pos.p, initial_positions, sizeof(double) * 2 * (size_t)pl->N,
                               cudaMemcpyHostToDevice, s));
    pl->pos_is_staged = true;
  } else {
    LFR_TRY(pl->pos_init.reserve(sizeof(double) * 2 * N));
    if (initial_positions && pl->N)
      LFR_CUDA(cudaMemcpyAsync(pl->pos_init.p, initial_positions, sizeof(double) * 2 * (size_t)pl->N,
                               cudaMemcpyHostToDevice, s));
    else
      LFR_CUDA(cudaMemsetAsync(pl->pos_init.p, 0, sizeof(double) * 2 * N, s));
    pl->pos_is_staged = false;
  }
  LFR_TRY(build_buckets(pl, p));  // host work overlaps the copies above
  if (zc_edges && pl->needs_hbm_edges) {
    // mixed schedule: the Cholesky-warp / CTA tiers read edges from global memory by index, so the
    // array goes to HBM after all — on its own stream, and only those tiers wait for it
    if (!pl->copy_stream) LFR_CUDA(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
    if (!pl->ev_edges) LFR_CUDA(cudaEventCreateWithFlags(&pl->ev_edges, cudaEventDisableTiming));
    LFR_TRY(upload(&pl->edges, p->edges, (size_t)p->n_edges, pl->copy_stream));
    LFR_CUDA(cudaEventRecord(pl->ev_edges, pl->copy_stream));
    pl->edges_in_hbm = true;
  }
  LFR_TRY(prepare_large(pl, p, s));
  LFR_TRY(upload(&pl->lists, pl->list_host.data(), pl->list_host.size(), s));
  if (pl->total_slots) {
    lfr::local_index_kernel<<<(pl->total_slots + 255) / 256, 256, 0, s>>>(
        pl->comp_ptr.as<uint32_t>(), pl->comp_nodes.as<uint32_t>(), pl->C, pl->total_slots,
        pl->local_of.as<uint32_t>());
    LFR_CUDA(cudaGetLastError());
  }
  // streams for concurrent bucket launches
  const int want = std::min<int>(kMaxStreams, std::max<int>(0, (int)pl->buckets.size() - 1 + (pl->n_large ? 1 : 0)));
  if (!pl->ev_fork) LFR_CUDA(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
  while (pl->n_streams < want) {
    LFR_CUDA(cudaStreamCreateWithFlags(&pl->streams[pl->n_streams], cudaStreamNonBlocking));
    LFR_CUDA(cudaEventCreateWithFlags(&pl->ev_join[pl->n_streams], cudaEventDisableTiming));
    ++pl->n_streams;
  }
  return LFR_OK;
}

int set_kernel_attrs() {
  static bool done_for_device[16] = {};  // per device, guarded by that device's workspace mutex / idempotent otherwise
  int dev = 0;
  LFR_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 16 && done_for_device[dev]) return LFR_OK;
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp2_kernel<4, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp2_kernel<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp2_kernel<4, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp2_kernel<4, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_warp2_kernel<1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_tile_kernel<64, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_tile_kernel<64, 48>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_tile_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  LFR_CUDA(cudaFuncSetAttribute(lfr::solve_tile_kernel<128, 80>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kMaxSmemPerBlock));
  if (dev >= 0 && dev < 16) done_for_device[dev] = true;
  return LFR_OK;
}

int launch_solve(lfr_plan* pl, cudaStream_t s) {
  LFR_TRY(set_kernel_attrs());
  if (pl->N && !pl->pos_is_staged)
    LFR_CUDA(cudaMemcpyAsync(pl->pos.p, pl->pos_init.p, sizeof(double) * 2 * (size_t)pl->N,
                             cudaMemcpyDeviceToDevice, s));
  pl->pos_is_staged = false;  // a second launch on the same plan needs the reset again
  const lfr::DevProblem P_hbm = pl->dev();  // edges read from the HBM copy by global index
  lfr::DevProblem P_stage = P_hbm;          // edges pulled once into shared memory, possibly from the caller's pinned buffer
  if (pl->zc_edges) P_stage.edges = pl->zc_edges;
  const bool hbm_edges_on_copy_stream = pl->zc_edges && pl->edges_in_hbm;
  const int nb = (int)pl->buckets.size();
  const int n_side = std::max(0, nb - 1) + ((pl->n_large && nb > 0) ? 1 : 0);
  if (n_side > 0) LFR_CUDA(cudaEventRecord(pl->ev_fork, s));
  int side = 0;
  if (pl->n_large) {  // the CTA tier first: its components are the longest
    cudaStream_t bs = s;
    if (nb > 0) {
      bs = pl->streams[side++ % pl->n_streams];
      LFR_CUDA(cudaStreamWaitEvent(bs, pl->ev_fork, 0));
    }
    lfr::CtaArrays A;
    A.eidx = pl->L_eidx.as<uint32_t>();
    A.meta = pl->L_meta.as<uint32_t>();
    A.inlist = pl->L_inlist.as<uint32_t>();
    A.twin = pl->L_twin.as<uint32_t>();
    A.fdst = pl->L_fdst.as<int32_t>();
    A.bmat = pl->L_bmat.as<double>();
    A.scr = pl->L_scr.as<double>();
    A.q = pl->L_q.as<double>();
    A.node = pl->L_node.as<uint32_t>();
    A.outptr = pl->L_outptr.as<uint32_t>();
    A.inptr = pl->L_inptr.as<uint32_t>();
    A.freeof = pl->L_freeof.as<int32_t>();
    A.x = pl->L_x.as<double>();
    A.xc = pl->L_xc.as<double>();
    A.lof = pl->L_lof.as<uint32_t>();
    A.vec = pl->L_vec.as<double>();
    A.total_free = pl->L_total_free;
    if (hbm_edges_on_copy_stream) LFR_CUDA(cudaStreamWaitEvent(bs, pl->ev_edges, 0));
    // dynamic shared memory for the CG vectors of the largest component of this launch (13 doubles per free node)
    const size_t cg_bytes = std::min<size_t>(200 * 1024, (size_t)pl->L_max_free * 13 * sizeof(double));
    lfr::solve_cta_kernel<<<pl->n_large, lfr::kCtaThreads, cg_bytes, bs>>>(P_hbm, pl->K, A, pl->L_comps.as<lfr::CtaComp>(),
                                                                          (unsigned)(cg_bytes / sizeof(double)));
    LFR_CUDA(cudaGetLastError());
  }
  for (int i = 0; i < nb; ++i) {
    const Bucket& b = pl->buckets[i];
    // bucket 0 runs on the caller's stream, the others on side streams forked from it
    cudaStream_t bs = s;
    if (i > 0) {
      bs = pl->streams[side++ % pl->n_streams];
      LFR_CUDA(cudaStreamWaitEvent(bs, pl->ev_fork, 0));
    }
    lfr::WarpBucket wb;
    wb.list = pl->lists.as<uint32_t>() + b.offset;
    wb.n = b.n;
    wb.emax = b.emax;
    wb.ncmax = b.ncmax;
    wb.n2max = b.n2max;
    wb.smem_per_warp = b.smem_per_warp;
    const size_t smem = (size_t)b.smem_per_warp * b.warps;
    const unsigned grid = (b.n + b.warps - 1) / b.warps;
    const lfr::DevProblem& P = b.stages_edges() ? P_stage : P_hbm;
    if (!b.stages_edges() && hbm_edges_on_copy_stream) LFR_CUDA(cudaStreamWaitEvent(bs, pl->ev_edges, 0));
    if (b.variant == 132)
      lfr::solve_tile_kernel<64, 32><<<b.n, 64, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 48)
      lfr::solve_tile_kernel<64, 48><<<b.n, 64, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 64)
      lfr::solve_tile_kernel<64, 64><<<b.n, 64, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 80)
      lfr::solve_tile_kernel<128, 80><<<b.n, 128, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 8)
      lfr::solve_warp2_kernel<4, 8><<<grid, 128, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 16)
      lfr::solve_warp2_kernel<4, 16><<<grid, 128, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 24 && b.warps == 4)
      lfr::solve_warp2_kernel<4, 24><<<grid, 128, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 32 && b.warps == 4)
      lfr::solve_warp2_kernel<4, 32><<<grid, 128, smem, bs>>>(P, pl->K, wb);
    else if (b.variant == 32)
      lfr::solve_warp2_kernel<1, 32><<<grid, 32, smem, bs>>>(P, pl->K, wb);
    else if (b.warps == 4)
      lfr::solve_warp_kernel<4><<<grid, 128, smem, bs>>>(P, pl->K, wb);
    else
      lfr::solve_warp_kernel<1><<<grid, 32, smem, bs>>>(P, pl->K, wb);
    LFR_CUDA(cudaGetLastError());
  }
  for (int k = 0; k < std::min(side, pl->n_streams); ++k) {  // join the side streams
    LFR_CUDA(cudaEventRecord(pl->ev_join[k], pl->streams[k]));
    LFR_CUDA(cudaStreamWaitEvent(s, pl->ev_join[k], 0));
  }
  return LFR_OK;
}

int download(lfr_plan* pl, cudaStream_t s, double* positions, lfr_stats* st) {
  const size_t nb = pl->stats_bytes();
  if (pl->h_stage_cap < nb) {
    if (pl->h_stage) cudaFreeHost(pl->h_stage);
    pl->h_stage = nullptr;
    pl->h_stage_cap = 0;
    LFR_CUDA(cudaMallocHost(&pl->h_stage, nb + nb / 4));
    pl->h_stage_cap = nb + nb / 4;
  }
  LFR_CUDA(cudaMemcpyAsync(pl->h_stage, pl->stats.p, nb, cudaMemcpyDeviceToHost, s));
  LFR_CUDA(cudaStreamSynchronize(s));
  const size_t Cp = pl->Cp;
  const double* h_cost0 = static_cast<const double*>(pl->h_stage);
  const double* h_cost1 = h_cost0 + Cp;
  const int32_t* h_iter = reinterpret_cast<const int32_t*>(h_cost1 + Cp);
  const int32_t* h_term = h_iter + Cp;
  const uint32_t* h_ls = reinterpret_cast<const uint32_t*>(h_term + Cp);
  const uint32_t* h_kept = h_ls + Cp;
  const int err = *reinterpret_cast<const int*>(h_kept + Cp);
  // the error flag is checked BEFORE anything is copied into the caller's position array (with
  // zero-copy write-back the components that did solve have already been written: on LFR_EINVAL
  // the array holds a mixture of start values and results)
  if (err) return fail(LFR_EINVAL, "edge with dst out of range or a self edge (found while staging edges on the device)");
  if (positions && pl->N && !pl->zc_positions) {
    LFR_CUDA(cudaMemcpyAsync(positions, pl->pos.p, sizeof(double) * 2 * (size_t)pl->N, cudaMemcpyDeviceToHost, s));
    LFR_CUDA(cudaStreamSynchronize(s));
  }
  if (st) {
    uint64_t ti = 0, tl = 0;
    for (uint32_t c = 0; c < pl->C; ++c) {
      ti += (uint64_t)h_iter[c];
      tl += h_ls[c];
    }
    if (pl->C) {
      if (st->iterations) std::memcpy(st->iterations, h_iter, sizeof(int32_t) * pl->C);
      if (st->termination) std::memcpy(st->termination, h_term, sizeof(int32_t) * pl->C);
      if (st->initial_cost) std::memcpy(st->initial_cost, h_cost0, sizeof(double) * pl->C);
      if (st->final_cost) std::memcpy(st->final_cost, h_cost1, sizeof(double) * pl->C);
    }
    st->total_iterations = ti;
    st->total_line_search_steps = tl;
    st->n_solved = pl->n_solved;
    st->n_kernel_launches = (uint32_t)pl->buckets.size() + (pl->n_large ? 1u : 0u);
  }
  return LFR_OK;
}

int select_device(const lfr_options& o) {
  int n_dev = 0;
  LFR_CUDA(cudaGetDeviceCount(&n_dev));
  if (n_dev == 0) return fail(LFR_ENODEV, "no CUDA device");
  if (o.device < 0 || o.device >= n_dev) return fail(LFR_ENODEV, "device ordinal out of range");
  LFR_CUDA(cudaSetDevice(o.device));
  return LFR_OK;
}

// lfr_solve()'s workspace: per device (shared by all host threads, guarded by g_ws_mutex), one cached
// plan (grow-only device buffers), one non-blocking stream and the events that time the call — all
// created with that device current.  Kept until lfr_shutdown() or process exit (no destructor calls
// into CUDA during teardown).
struct DeviceWorkspace {
  lfr_plan* plan = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {};
  bool ready = false;
};
constexpr int kMaxDevices = 16;
DeviceWorkspace g_ws[kMaxDevices];
std::mutex g_ws_mutex[kMaxDevices];  // one solve at a time per device workspace; distinct devices run concurrently

int ensure_workspace(int device) {
  DeviceWorkspace& ws = g_ws[device];
  if (ws.ready) return LFR_OK;
  LFR_CUDA(cudaStreamCreateWithFlags(&ws.stream, cudaStreamNonBlocking));
  for (int i = 0; i < 4; ++i) LFR_CUDA(cudaEventCreate(&ws.ev[i]));
  ws.plan = new lfr_plan();
  ws.plan->device = device;
  ws.ready = true;
  return LFR_OK;
}

// One device's part of a solve: uploads, launches, results.  `slot_owner` / `owner_id` restrict the
// plan to the dispatch slots this device owns (nullptr = all).  Caller holds g_ws_mutex[device].
int solve_on_device(const lfr_problem* p, const lfr_options& o, double* positions, lfr_stats* st,
                    const uint8_t* slot_owner, uint8_t owner_id, double* pos_scratch) {
  LFR_TRY(select_device(o));
  LFR_TRY(ensure_workspace(o.device));
  DeviceWorkspace& ws = g_ws[o.device];
  lfr_plan* pl = ws.plan;
  cudaStream_t s = ws.stream;
  pl->slot_owner = slot_owner;
  pl->owner_id = owner_id;
  // page-locked caller buffers are used in place (see include/lfr.h); pageable ones go through HBM
  const bool zero_copy = !(o.debug_flags & LFR_DBG_NO_ZERO_COPY);
  const float4* zc_edges = (zero_copy && p->n_edges) ? static_cast<const float4*>(device_view_of_pinned(p->edges)) : nullptr;
  double* zc_positions = (zero_copy && p->n_nodes) ? static_cast<double*>(device_view_of_pinned(positions)) : nullptr;
  LFR_CUDA(cudaEventRecord(ws.ev[0], s));
  LFR_TRY(fill_plan(pl, p, o, positions, s, /*stage_positions_directly=*/true, zc_edges, zc_positions));
  LFR_CUDA(cudaEventRecord(ws.ev[1], s));
  LFR_TRY(launch_solve(pl, s));
  LFR_CUDA(cudaEventRecord(ws.ev[2], s));
  // several devices, pageable positions: each device returns its copy into its own scratch array
  // and the caller merges the entries of the components it owns
  int rc = download(pl, s, (slot_owner && !zc_positions) ? pos_scratch : positions, st);
  if (pl->copy_stream && pl->zc_edges && pl->edges_in_hbm) {
    // the bulk copy reads the caller's buffer: it must be over before the call returns
    cudaError_t e = cudaStreamSynchronize(pl->copy_stream);
    if (e != cudaSuccess && rc == LFR_OK) rc = fail(cuda_code(e), cudaGetErrorString(e));
  }
  pl->slot_owner = nullptr;
  if (rc) return rc;
  LFR_CUDA(cudaEventRecord(ws.ev[3], s));
  LFR_CUDA(cudaEventSynchronize(ws.ev[3]));
  if (st) {
    float a = 0, b = 0, c = 0;
    cudaEventElapsedTime(&a, ws.ev[0], ws.ev[1]);
    cudaEventElapsedTime(&b, ws.ev[1], ws.ev[2]);
    cudaEventElapsedTime(&c, ws.ev[2], ws.ev[3]);
    st->h2d_ms = a;     // uploads of the arrays that go through HBM (+ schedule)
    st->kernel_ms = b;  // the solve kernels (with zero-copy: including their PCIe pulls)
    st->d2h_ms = c;
    st->total_ms = (double)a + b + c;
  }
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
  LFR_TRY(validate(p));
  lfr_options o;
  if (opt) o = *opt; else lfr_options_default(&o);
  LFR_TRY(select_device(o));
  lfr_plan* pl = new lfr_plan();
  pl->device = o.device;
  int rc = fill_plan(pl, p, o, initial_positions, 0);
  if (rc == LFR_OK) {
    cudaError_t e = cudaStreamSynchronize(0);
    if (e != cudaSuccess) rc = fail(cuda_code(e), cudaGetErrorString(e));
  }
  if (rc) {
    free_plan(pl);
    return rc;
  }
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
  return pl ? (int)pl->buckets.size() + (pl->n_large ? 1 : 0) : 0;
}

int lfr_plan_download(lfr_plan* pl, void* stream, double* positions, lfr_stats* st) {
  if (!pl) return fail(LFR_EINVAL, "plan is NULL");
  LFR_CUDA(cudaSetDevice(pl->device));
  return download(pl, (cudaStream_t)stream, positions, st);
}

int lfr_plan_traffic(lfr_plan* pl, void* stream, uint64_t* algorithmic_bytes, uint64_t* one_pass_bytes) {
  if (!pl) return fail(LFR_EINVAL, "plan is NULL");
  LFR_CUDA(cudaSetDevice(pl->device));
  cudaStream_t s = (cudaStream_t)stream;
  std::vector<unsigned char> blk(pl->stats_bytes());
  LFR_CUDA(cudaMemcpyAsync(blk.data(), pl->stats.p, blk.size(), cudaMemcpyDeviceToHost, s));
  LFR_CUDA(cudaStreamSynchronize(s));
  const int32_t* it = reinterpret_cast<const int32_t*>(blk.data() + 16 * (size_t)pl->Cp);
  const uint32_t* kept = reinterpret_cast<const uint32_t*>(blk.data() + 16 * (size_t)pl->Cp + 12 * (size_t)pl->Cp);
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

/* debug (LFR_PROFILE=1): per-slot cycle counters, 8 x uint64 each */
int lfr_debug_plan_cycles(lfr_plan* pl, unsigned long long* out) {
  if (!pl || !pl->profile) return fail(LFR_EINVAL, "plan was not created with LFR_PROFILE=1");
  LFR_CUDA(cudaSetDevice(pl->device));
  LFR_CUDA(cudaMemcpy(out, pl->cycles.p, sizeof(unsigned long long) * 8 * (size_t)pl->C, cudaMemcpyDeviceToHost));
  return LFR_OK;
}

/* debug (LFR_DBG_PROFILE): counters of the last lfr_solve() of this thread on `device`:
   cycles [8 per slot], times [2 per slot] = %globaltimer ns at start / end of each component */
int lfr_debug_last_solve_profile(int device, unsigned long long* cycles, unsigned long long* times) {
  if (device < 0 || device >= kMaxDevices || !g_ws[device].ready || !g_ws[device].plan->profile)
    return fail(LFR_EINVAL, "no profiled lfr_solve() on this device / thread");
  lfr_plan* pl = g_ws[device].plan;
  LFR_CUDA(cudaSetDevice(device));
  if (cycles) LFR_CUDA(cudaMemcpy(cycles, pl->cycles.p, sizeof(unsigned long long) * 8 * (size_t)pl->C, cudaMemcpyDeviceToHost));
  if (times) LFR_CUDA(cudaMemcpy(times, pl->times.p, sizeof(unsigned long long) * 2 * (size_t)pl->C, cudaMemcpyDeviceToHost));
  return LFR_OK;
}

#ifdef LFR_POLY_PROF
/* diagnostic build only: cycle split of the line-search polynomial (lfr_math.cuh) */
int lfr_debug_poly_prof(unsigned long long* out16, int reset) {
  LFR_CUDA(cudaMemcpyFromSymbol(out16, lfr::g_poly_prof, sizeof(unsigned long long) * 16));
  if (reset) {
    unsigned long long z[16] = {0};
    LFR_CUDA(cudaMemcpyToSymbol(lfr::g_poly_prof, z, sizeof(z)));
  }
  return LFR_OK;
}
#endif

void lfr_plan_destroy(lfr_plan* pl) {
  if (!pl) return;
  cudaSetDevice(pl->device);
  free_plan(pl);
}

int lfr_solve(const lfr_problem* p, const lfr_options* opt, double* positions, lfr_stats* st) {
  LFR_TRY(validate(p));
  if (p->n_nodes && !positions) return fail(LFR_EINVAL, "positions is NULL");
  lfr_options o;
  if (opt) o = *opt; else lfr_options_default(&o);
  if (o.device < 0 || o.device >= kMaxDevices) return fail(LFR_EUNSUPPORTED, "device ordinal out of [0, 16)");
  std::lock_guard<std::mutex> lock(g_ws_mutex[o.device]);
  return solve_on_device(p, o, positions, st, nullptr, 0, nullptr);
}

void lfr_shutdown(void) {
  // release the cached per-device workspaces of lfr_solve() / lfr_solve_multi() (device buffers,
  // pinned staging, streams, events); the next call re-creates what it needs
  for (int d = 0; d < kMaxDevices; ++d) {
    std::lock_guard<std::mutex> lock(g_ws_mutex[d]);
    DeviceWorkspace& ws = g_ws[d];
    if (!ws.ready) continue;
    if (cudaSetDevice(d) == cudaSuccess) {
      free_plan(ws.plan);
      for (int i = 0; i < 4; ++i)
        if (ws.ev[i]) cudaEventDestroy(ws.ev[i]);
      if (ws.stream) cudaStreamDestroy(ws.stream);
    }
    ws = DeviceWorkspace();
  }
}

int lfr_solve_multi(const lfr_problem* p, const lfr_options* opt, const int32_t* devices, int32_t n_devices,
                    double* positions, lfr_stats* st, lfr_multi_info* info) {
  LFR_TRY(validate(p));
  if (p->n_nodes && !positions) return fail(LFR_EINVAL, "positions is NULL");
  if (!devices || n_devices < 1 || n_devices > kMaxDevices) return fail(LFR_EINVAL, "devices: need 1..16 device ordinals");
  for (int d = 0; d < n_devices; ++d) {
    if (devices[d] < 0 || devices[d] >= kMaxDevices) return fail(LFR_EUNSUPPORTED, "device ordinal out of [0, 16)");
    for (int e = 0; e < d; ++e)
      if (devices[e] == devices[d]) return fail(LFR_EINVAL, "devices: duplicate ordinal");
  }
  lfr_options o0;
  if (opt) o0 = *opt; else lfr_options_default(&o0);
  const uint32_t C = p->n_components;
  // ---- LPT packing of the dispatch slots by directed-edge count, largest first (the reference's
  // largest-first queue, solve.cc:599-604, spread over devices); deterministic
  std::vector<uint64_t> weight(C, 0);
  const uint32_t total_slots = C ? p->comp_ptr[C] : 0;
  for (uint32_t c = 0; c < C; ++c) {
    const uint32_t beg = p->comp_ptr[c], end = p->comp_ptr[c + 1];
    if (end < beg || end > total_slots) return fail(LFR_EINVAL, "comp_ptr not monotone");
    if (end - beg <= 1) continue;
    for (uint32_t i = beg; i < end; ++i) {
      const uint32_t v = p->comp_nodes[i];
      if (v >= p->n_nodes) return fail(LFR_EINVAL, "comp_nodes out of range");
      weight[c] += p->row_ptr[v + 1] - p->row_ptr[v];
    }
  }
  std::vector<uint32_t> order(C);
  for (uint32_t c = 0; c < C; ++c) order[c] = c;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
  std::vector<uint8_t> owner(C, 0);
  std::vector<uint64_t> load(n_devices, 0);
  std::vector<uint32_t> n_slots(n_devices, 0);
  for (uint32_t c : order) {
    int best = 0;
    for (int d = 1; d < n_devices; ++d)
      if (load[d] < load[best]) best = d;
    owner[c] = (uint8_t)best;
    load[best] += weight[c];
    if (weight[c]) ++n_slots[best];
  }
  // ---- one host thread per device: upload the (small) per-node arrays, launch, collect.  With
  // page-locked caller buffers every device pulls only ITS components' edge records from the shared
  // host array and writes its results straight into `positions` (disjoint entries): the edge data is
  // partitioned without ever being copied, and no collective is needed.
  std::vector<lfr_stats> dst(n_devices);
  std::vector<std::vector<int32_t>> d_iter(n_devices), d_term(n_devices);
  std::vector<std::vector<double>> d_c0(n_devices), d_c1(n_devices), d_pos(n_devices);
  std::vector<int> rcs(n_devices, LFR_OK);
  std::vector<std::string> errs(n_devices);
  const bool pos_pinned = !(o0.debug_flags & LFR_DBG_NO_ZERO_COPY) && p->n_nodes && device_view_of_pinned(positions) != nullptr;
  auto work = [&](int d) {
    lfr_options o = o0;
    o.device = devices[d];
    std::memset(&dst[d], 0, sizeof(lfr_stats));
    d_iter[d].assign(C, 0);
    d_term[d].assign(C, 0);
    d_c0[d].assign(C, 0.0);
    d_c1[d].assign(C, 0.0);
    dst[d].iterations = d_iter[d].data();
    dst[d].termination = d_term[d].data();
    dst[d].initial_cost = d_c0[d].data();
    dst[d].final_cost = d_c1[d].data();
    if (!pos_pinned) d_pos[d].assign(2 * (size_t)p->n_nodes, 0.0);
    std::lock_guard<std::mutex> lock(g_ws_mutex[o.device]);
    rcs[d] = solve_on_device(p, o, positions, &dst[d], owner.data(), (uint8_t)d, pos_pinned ? nullptr : d_pos[d].data());
    if (rcs[d]) errs[d] = g_last_error;
  };
  if (n_devices == 1) {
    work(0);
  } else {
    std::vector<std::thread> threads;
    for (int d = 0; d < n_devices; ++d) threads.emplace_back(work, d);
    for (auto& t : threads) t.join();
  }
  for (int d = 0; d < n_devices; ++d)
    if (rcs[d]) return fail(rcs[d], "device " + std::to_string(devices[d]) + ": " + errs[d]);
  // ---- merge by owner
  if (!pos_pinned) {
    for (uint32_t c = 0; c < C; ++c) {
      const uint32_t beg = p->comp_ptr[c], end = p->comp_ptr[c + 1];
      if (end - beg <= 1) continue;
      const double* src = d_pos[owner[c]].data();
      for (uint32_t i = beg; i < end; ++i) {
        const size_t v = p->comp_nodes[i];
        positions[2 * v] = src[2 * v];
        positions[2 * v + 1] = src[2 * v + 1];
      }
    }
  }
  if (st) {
    uint64_t ti = 0, tl = 0;
    uint32_t ns = 0, nk = 0;
    double h = 0, k = 0, dd = 0, tt = 0;
    for (int d = 0; d < n_devices; ++d) {
      ti += dst[d].total_iterations;
      tl += dst[d].total_line_search_steps;
      ns += dst[d].n_solved;
      nk += dst[d].n_kernel_launches;
      h = std::max(h, dst[d].h2d_ms);
      k = std::max(k, dst[d].kernel_ms);
      dd = std::max(dd, dst[d].d2h_ms);
      tt = std::max(tt, dst[d].total_ms);
    }
    for (uint32_t c = 0; c < C; ++c) {
      const int d = owner[c];
      if (st->iterations) st->iterations[c] = d_iter[d][c];
      if (st->termination) st->termination[c] = d_term[d][c];
      if (st->initial_cost) st->initial_cost[c] = d_c0[d][c];
      if (st->final_cost) st->final_cost[c] = d_c1[d][c];
    }
    st->total_iterations = ti;
    st->total_line_search_steps = tl;
    st->n_solved = ns;
    st->n_kernel_launches = nk;
    st->h2d_ms = h;       // maxima over the devices (they run concurrently)
    st->kernel_ms = k;
    st->d2h_ms = dd;
    st->total_ms = tt;
  }
  if (info) {
    for (int d = 0; d < n_devices && d < 16; ++d) {
      info->kernel_ms[d] = dst[d].kernel_ms;
      info->total_ms[d] = dst[d].total_ms;
      info->n_slots[d] = n_slots[d];
      info->n_edges[d] = load[d];
    }
    info->zero_copy = (pos_pinned && p->n_edges && device_view_of_pinned(p->edges) != nullptr) ? 1 : 0;
  }
  return LFR_OK;
}

int lfr_debug_edge_eval(const lfr_edge* edges, const uint8_t* kind, uint64_t n, const double* xs,
                        const double* xd, const lfr_options* opt, double* r, double* jac, double* rho) {
  lfr_options o;
  if (opt) o = *opt; else lfr_options_default(&o);
  LFR_TRY(select_device(o));
  if (n == 0) return LFR_OK;
  DevBuf d_e, d_k, d_xs, d_xd, d_r, d_j, d_rho;
  int rc = LFR_OK;
  auto run = [&]() -> int {
    LFR_TRY(upload(&d_e, edges, n, 0));
    LFR_TRY(upload(&d_k, kind, n, 0));
    LFR_TRY(upload(&d_xs, xs, 2 * n, 0));
    LFR_TRY(upload(&d_xd, xd, 2 * n, 0));
    LFR_TRY(d_r.reserve(n * 16));
    LFR_TRY(d_j.reserve(n * 32));
    LFR_TRY(d_rho.reserve(n * 24));
    lfr::edge_eval_kernel<<<(unsigned)((n + 127) / 128), 128>>>(d_e.as<float4>(), d_k.as<uint8_t>(), n,
                                                               d_xs.as<double>(), d_xd.as<double>(),
                                                               make_consts(o), d_r.as<double>(), d_j.as<double>(),
                                                               d_rho.as<double>());
    LFR_CUDA(cudaGetLastError());
    LFR_CUDA(cudaMemcpy(r, d_r.p, n * 16, cudaMemcpyDeviceToHost));
    LFR_CUDA(cudaMemcpy(jac, d_j.p, n * 32, cudaMemcpyDeviceToHost));
    LFR_CUDA(cudaMemcpy(rho, d_rho.p, n * 24, cudaMemcpyDeviceToHost));
    return LFR_OK;
  };
  rc = run();
  d_e.release(); d_k.release(); d_xs.release(); d_xd.release(); d_r.release(); d_j.release(); d_rho.release();
  return rc;
}

/* test hook: real roots in [lo, hi] of n quartics (5 coefficients each, highest degree first,
   leading coefficient != 0) by the line search's root finders; use_grid = 1 -> Budan-Fourier grid
   isolation (the production route), 0 -> derivative recursion */
int lfr_debug_quartic_roots(const double* coef, const double* lohi, uint64_t n, int use_grid, double* roots,
                            int* counts) {
  lfr_options o;
  lfr_options_default(&o);
  LFR_TRY(select_device(o));
  if (n == 0) return LFR_OK;
  DevBuf d_c, d_l, d_r, d_n;
  auto run = [&]() -> int {
    LFR_TRY(upload(&d_c, coef, 5 * n, 0));
    LFR_TRY(upload(&d_l, lohi, 2 * n, 0));
    LFR_TRY(d_r.reserve(n * 32));
    LFR_TRY(d_n.reserve(n * 4));
    lfr::quartic_roots_kernel<<<(unsigned)((n + 3) / 4), 128>>>(d_c.as<double>(), d_l.as<double>(), (int)n, use_grid,
                                                                d_r.as<double>(), d_n.as<int>());
    LFR_CUDA(cudaGetLastError());
    LFR_CUDA(cudaMemcpy(roots, d_r.p, n * 32, cudaMemcpyDeviceToHost));
    LFR_CUDA(cudaMemcpy(counts, d_n.p, n * 4, cudaMemcpyDeviceToHost));
    return LFR_OK;
  };
  const int rc = run();
  d_c.release(); d_l.release(); d_r.release(); d_n.release();
  return rc;
}

/* test hook: the line search's interpolating-polynomial minimiser on n cases of 11 doubles
   {f0, g0, x1, f1, g1, three, x2, f2, g2, lo, hi}; out[n] = selected step size */
int lfr_debug_ls_minimizer(const double* cases, uint64_t n, double* out) {
  lfr_options o;
  lfr_options_default(&o);
  LFR_TRY(select_device(o));
  if (n == 0) return LFR_OK;
  DevBuf d_i, d_o;
  auto run = [&]() -> int {
    LFR_TRY(upload(&d_i, cases, 11 * n, 0));
    LFR_TRY(d_o.reserve(n * 8));
    lfr::ls_minimizer_kernel<<<(unsigned)((n + 3) / 4), 128>>>(d_i.as<double>(), (int)n, d_o.as<double>());
    LFR_CUDA(cudaGetLastError());
    LFR_CUDA(cudaMemcpy(out, d_o.p, n * 8, cudaMemcpyDeviceToHost));
    return LFR_OK;
  };
  const int rc = run();
  d_i.release(); d_o.release();
  return rc;
}

}  // extern "C"
