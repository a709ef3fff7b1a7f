// lfr_host.cc — native host graph stage (include/lfr_host.h): solve.cc:438-606
// from flat match arrays to the arrays of lfr_problem.  Mirrors
// local-feature-refinement_b200/graph.py statement for statement where order
// matters (tie-breaks, accumulation order, the deterministic 2-way cut), and is
// tested to produce identical arrays.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <map>
#include <numeric>
#include <unordered_map>
#include <vector>

#include "../../include/lfr_host.h"

struct lfr_host_stage {
  uint32_t N = 0, T = 0, C = 0;
  uint64_t E = 0;
  std::vector<uint32_t> row_ptr, track, comp, comp_ptr, comp_nodes, comp_order, node_image, node_feat;
  std::vector<uint8_t> is_root;
  std::vector<lfr_edge> edges;
};

namespace {

using Clock = std::chrono::steady_clock;
double ms_since(Clock::time_point t0) { return std::chrono::duration<double, std::milli>(Clock::now() - t0).count(); }

// ---- connected components, labelled in order of the lowest member index (solve.cc:291-300)
std::vector<uint32_t> connected_components(uint32_t n, const std::vector<uint32_t>& a, const std::vector<uint32_t>& b,
                                           const std::vector<uint8_t>* keep, uint32_t* n_out) {
  std::vector<uint32_t> parent(n);
  std::iota(parent.begin(), parent.end(), 0u);
  auto find = [&](uint32_t x) {
    while (parent[x] != x) {
      parent[x] = parent[parent[x]];
      x = parent[x];
    }
    return x;
  };
  for (size_t i = 0; i < a.size(); ++i) {
    if (keep && !(*keep)[i]) continue;
    const uint32_t ra = find(a[i]), rb = find(b[i]);
    if (ra != rb) parent[std::max(ra, rb)] = std::min(ra, rb);
  }
  std::vector<uint32_t> label(n), of_root(n, UINT32_MAX);
  uint32_t next = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t r = find(i);
    if (of_root[r] == UINT32_MAX) of_root[r] = next++;
    label[i] = of_root[r];
  }
  *n_out = next;
  return label;
}

typedef std::map<uint32_t, std::map<uint32_t, int64_t>> Adj;

std::vector<uint32_t> bfs_order(const Adj& adj, uint32_t start) {
  std::vector<uint32_t> order(1, start);
  std::map<uint32_t, bool> mark;
  mark[start] = true;
  for (size_t head = 0; head < order.size(); ++head) {
    const uint32_t u = order[head];
    for (const auto& kv : adj.at(u)) {  // std::map iterates neighbours in ascending order
      if (!mark.count(kv.first)) {
        mark[kv.first] = true;
        order.push_back(kv.first);
      }
    }
  }
  return order;
}

// graph.py::two_way_cut — deterministic stand-in for ComputeNormalizedMinGraphCut(edges, weights, 2)
std::map<uint32_t, int> two_way_cut(const Adj& adj, const std::vector<uint32_t>& node_weight) {
  std::vector<uint32_t> nodes;
  for (const auto& kv : adj) nodes.push_back(kv.first);  // ascending
  std::map<uint32_t, bool> seen;
  std::vector<std::vector<uint32_t>> pieces;
  for (uint32_t s : nodes) {
    if (seen.count(s)) continue;
    std::vector<uint32_t> comp(1, s);
    seen[s] = true;
    for (size_t head = 0; head < comp.size(); ++head) {
      const uint32_t u = comp[head];
      for (const auto& kv : adj.at(u))
        if (!seen.count(kv.first)) {
          seen[kv.first] = true;
          comp.push_back(kv.first);
        }
    }
    pieces.push_back(comp);
  }
  std::map<uint32_t, int> out;
  if (pieces.size() > 1) {
    std::vector<std::pair<int64_t, size_t>> key(pieces.size());
    for (size_t i = 0; i < pieces.size(); ++i) {
      int64_t w = 0;
      for (uint32_t x : pieces[i]) w += node_weight[x];
      key[i] = std::make_pair(w, i);
    }
    std::stable_sort(key.begin(), key.end(), [&](const std::pair<int64_t, size_t>& p, const std::pair<int64_t, size_t>& q) {
      if (p.first != q.first) return p.first > q.first;               // heaviest first
      return pieces[p.second][0] < pieces[q.second][0];               // then by first node
    });
    int64_t w[2] = {0, 0};
    for (const auto& k : key) {
      const int side = (w[0] <= w[1]) ? 0 : 1;
      w[side] += k.first;
      for (uint32_t x : pieces[k.second]) out[x] = side;
    }
    return out;
  }
  const uint32_t start = bfs_order(adj, nodes[0]).back();
  const std::vector<uint32_t> order = bfs_order(adj, start);
  int64_t total = 0, acc = 0;
  for (uint32_t x : nodes) total += node_weight[x];
  for (size_t i = 0; i < order.size(); ++i) {
    if (i > 0 && (acc * 2 >= total || i == order.size() - 1)) break;
    out[order[i]] = 0;
    acc += node_weight[order[i]];
  }
  for (uint32_t x : order)
    if (!out.count(x)) out[x] = 1;
  int cnt[2] = {0, 0};
  for (const auto& kv : out) ++cnt[kv.second];
  for (uint32_t x : order) {  // one refinement sweep
    const int s = out[x];
    if (cnt[s] <= 1) continue;
    int64_t inside = 0, outside = 0;
    for (const auto& kv : adj.at(x)) {
      if (out[kv.first] == s) inside += kv.second; else outside += kv.second;
    }
    if (outside > inside) {
      out[x] = 1 - s;
      --cnt[s];
      ++cnt[1 - s];
    }
  }
  return out;
}

struct MetaEdge {
  uint32_t a, b;
  int64_t w;
};

// graph.py::recursive_cut (solve.cc:185-250 as a work list)
void recursive_cut(const std::vector<MetaEdge>& edges0, const std::vector<uint32_t>& node_weight, uint32_t max_weight,
                   std::vector<std::vector<uint32_t>>* groups) {
  std::vector<std::vector<MetaEdge>> work(1, edges0);
  while (!work.empty()) {
    std::vector<MetaEdge> edges = std::move(work.back());
    work.pop_back();
    Adj adj;
    for (const MetaEdge& e : edges) {
      adj[e.a][e.b] += e.w;
      adj[e.b][e.a] += e.w;
    }
    const std::map<uint32_t, int> side = two_way_cut(adj, node_weight);
    for (int s = 0; s < 2; ++s) {
      std::vector<uint32_t> members;
      int64_t w = 0;
      for (const auto& kv : side)
        if (kv.second == s) {
          members.push_back(kv.first);  // ascending
          w += node_weight[kv.first];
        }
      if (members.empty()) continue;
      if (w <= (int64_t)max_weight) {
        groups->push_back(members);  // solve.cc:205-211
        continue;
      }
      std::vector<MetaEdge> sub;
      std::map<uint32_t, bool> covered;
      for (const MetaEdge& e : edges) {
        const auto ia = side.find(e.a), ib = side.find(e.b);
        if (ia->second == s && ib->second == s) {
          sub.push_back(e);
          covered[e.a] = true;
          covered[e.b] = true;
        }
      }
      if (!sub.empty()) {
        work.push_back(sub);
        for (uint32_t x : members)
          if (!covered.count(x)) groups->push_back(std::vector<uint32_t>(1, x));
      } else {
        for (uint32_t x : members) groups->push_back(std::vector<uint32_t>(1, x));
      }
    }
  }
}

}  // namespace

extern "C" {

int lfr_host_stage_create(const lfr_host_input* in, lfr_host_stage** out, lfr_host_sizes* sz) {
  if (!in || !out) return LFR_EINVAL;
  *out = nullptr;
  lfr_host_stage* hs = new lfr_host_stage();
  lfr_host_sizes S;
  std::memset(&S, 0, sizeof S);
  // ---- H1: node interning + directed edges (solve.cc:438-481) -----------------------------
  std::vector<uint64_t> kept_matches;  // indices of matches of non-skipped pairs, in order
  std::vector<uint32_t> m_img1, m_img2;
  std::vector<uint8_t> seen_img(in->n_images, 0);
  for (uint64_t p = 0; p < in->n_pairs; ++p) {
    if (in->pair_skip && in->pair_skip[p]) continue;
    if (in->pair_img1[p] >= in->n_images || in->pair_img2[p] >= in->n_images) {
      delete hs;
      return LFR_EINVAL;
    }
    seen_img[in->pair_img1[p]] = 1;
    seen_img[in->pair_img2[p]] = 1;
    for (uint64_t m = in->pair_ptr[p]; m < in->pair_ptr[p + 1]; ++m) {
      kept_matches.push_back(m);
      m_img1.push_back(in->pair_img1[p]);
      m_img2.push_back(in->pair_img2[p]);
    }
  }
  for (uint8_t v : seen_img) S.n_images_seen += v;
  const uint64_t M = kept_matches.size();
  std::vector<uint32_t> n1(M), n2(M);
  {
    std::unordered_map<uint64_t, uint32_t> ids;
    ids.reserve((size_t)(2 * M * 1.3) + 16);
    auto intern = [&](uint32_t img, uint32_t feat) {
      const uint64_t key = ((uint64_t)img << 32) | feat;
      auto it = ids.find(key);
      if (it != ids.end()) return it->second;
      const uint32_t id = (uint32_t)hs->node_image.size();
      ids.emplace(key, id);
      hs->node_image.push_back(img);
      hs->node_feat.push_back(feat);
      return id;
    };
    for (uint64_t k = 0; k < M; ++k) {
      const uint64_t m = kept_matches[k];
      n1[k] = intern(m_img1[k], in->feat1[m]);   // find_or_create_node: side 1 then side 2
      n2[k] = intern(m_img2[k], in->feat2[m]);
    }
  }
  const uint32_t N = (uint32_t)hs->node_image.size();
  hs->N = N;
  hs->E = 2 * M;
  hs->row_ptr.assign((size_t)N + 1, 0);
  for (uint64_t k = 0; k < M; ++k) {
    ++hs->row_ptr[n1[k] + 1];
    ++hs->row_ptr[n2[k] + 1];
  }
  for (uint32_t v = 0; v < N; ++v) hs->row_ptr[v + 1] += hs->row_ptr[v];
  hs->edges.resize(2 * M);
  {
    std::vector<uint32_t> fill(hs->row_ptr.begin(), hs->row_ptr.end() - (N ? 1 : 0));
    for (uint64_t k = 0; k < M; ++k) {  // add_edge order: n1->n2 (disp2) then n2->n1 (disp1), solve.cc:477-478
      const uint64_t m = kept_matches[k];
      lfr_edge& e1 = hs->edges[fill[n1[k]]++];
      std::memcpy(e1.flow, in->disp2 + 18 * m, 18 * sizeof(float));
      e1.sim = in->sim[m];
      e1.dst = n2[k];
      lfr_edge& e2 = hs->edges[fill[n2[k]]++];
      std::memcpy(e2.flow, in->disp1 + 18 * m, 18 * sizeof(float));
      e2.sim = in->sim[m];
      e2.dst = n1[k];
    }
  }
  if (N == 0) {
    hs->comp_ptr.assign(1, 0);
    *out = hs;
    if (sz) *sz = S;
    return LFR_OK;
  }
  const auto t_tracks = Clock::now();
  // ---- H2: constrained Kruskal (solve.cc:489-541) ------------------------------------------
  std::vector<uint64_t> order(M);
  std::iota(order.begin(), order.end(), (uint64_t)0);
  std::sort(order.begin(), order.end(), [&](uint64_t x, uint64_t y) {  // sort + reverse on (sim, n1, n2)
    const float sx = in->sim[kept_matches[x]], sy = in->sim[kept_matches[y]];
    if (sx != sy) return sx > sy;
    if (n1[x] != n1[y]) return n1[x] > n1[y];
    if (n2[x] != n2[y]) return n2[x] > n2[y];
    return x > y;
  });
  std::vector<int32_t> parent(N, -1);
  std::vector<std::vector<uint16_t>> imgs(N);
  for (uint32_t v = 0; v < N; ++v) imgs[v].assign(1, (uint16_t)hs->node_image[v]);
  if (in->n_images > 65535) {
    delete hs;
    return LFR_EUNSUPPORTED;
  }
  auto find = [&](uint32_t x) {
    uint32_t r = x;
    while (parent[r] != -1) r = (uint32_t)parent[r];
    while (parent[x] != -1) {  // path compression (solve.cc:74-76)
      const uint32_t nx = (uint32_t)parent[x];
      parent[x] = (int32_t)r;
      x = nx;
    }
    return r;
  };
  std::vector<uint16_t> merged;
  for (uint64_t k : order) {
    const uint32_t r1 = find(n1[k]), r2 = find(n2[k]);
    if (r1 == r2) continue;
    const std::vector<uint16_t>&a = imgs[r1], &b = imgs[r2];
    bool clash = false;  // std::set_intersection non-empty (solve.cc:507-511)
    for (size_t i = 0, j = 0; i < a.size() && j < b.size();) {
      if (a[i] == b[j]) { clash = true; break; }
      if (a[i] < b[j]) ++i; else ++j;
    }
    if (clash) continue;
    merged.resize(a.size() + b.size());
    std::merge(a.begin(), a.end(), b.begin(), b.end(), merged.begin());
    if (a.size() < b.size()) {  // solve.cc:513-521
      parent[r1] = (int32_t)r2;
      imgs[r2] = merged;
      std::vector<uint16_t>().swap(imgs[r1]);
    } else {
      parent[r2] = (int32_t)r1;
      imgs[r1] = merged;
      std::vector<uint16_t>().swap(imgs[r2]);
    }
  }
  hs->track.assign(N, 0);
  uint32_t T = 0;
  for (uint32_t v = 0; v < N; ++v)
    if (parent[v] == -1) hs->track[v] = T++;
  for (uint32_t v = 0; v < N; ++v)
    if (parent[v] != -1) hs->track[v] = hs->track[find(v)];
  hs->T = T;
  std::vector<uint32_t> nodes_in_track(T, 0);
  for (uint32_t v = 0; v < N; ++v) ++nodes_in_track[hs->track[v]];
  S.max_track_size = *std::max_element(nodes_in_track.begin(), nodes_in_track.end());
  std::vector<std::vector<uint16_t>>().swap(imgs);
  // ---- H3: roots (solve.cc:552-582) -----------------------------------------------------------
  hs->is_root.assign(N, 0);
  {
    std::vector<double> best_score(T, -1.0);
    std::vector<uint32_t> best_node(T, 0);
    std::vector<uint8_t> has(T, 0);
    for (uint32_t v = 0; v < N; ++v) {
      double score = 0.0;
      for (uint32_t e = hs->row_ptr[v]; e < hs->row_ptr[v + 1]; ++e)
        if (hs->track[v] == hs->track[hs->edges[e].dst]) score += (double)hs->edges[e].sim;
      const uint32_t t = hs->track[v];
      if (!has[t] || score > best_score[t] || (score == best_score[t] && v > best_node[t])) {
        has[t] = 1;
        best_score[t] = score;
        best_node[t] = v;
      }
    }
    for (uint32_t t = 0; t < T; ++t) hs->is_root[best_node[t]] = 1;
  }
  S.tracks_ms = ms_since(t_tracks);
  const auto t_cut = Clock::now();
  // ---- H4: meta-graph, connected components, size-capped cut (solve.cc:252-373) -------------
  std::vector<uint32_t> ma, mb;
  std::vector<double> wsum;
  {
    std::unordered_map<uint64_t, uint32_t> slot;
    std::vector<uint64_t> keys;
    std::vector<double> sums;
    for (uint32_t v = 0; v < N; ++v) {
      const uint32_t ts = hs->track[v];
      for (uint32_t e = hs->row_ptr[v]; e < hs->row_ptr[v + 1]; ++e) {
        const uint32_t tt = hs->track[hs->edges[e].dst];
        if (ts == tt) continue;
        const uint64_t key = (uint64_t)ts * T + tt;
        auto it = slot.find(key);
        if (it == slot.end()) {
          slot.emplace(key, (uint32_t)keys.size());
          keys.push_back(key);
          sums.push_back((double)hs->edges[e].sim);
        } else {
          sums[it->second] += (double)hs->edges[e].sim;  // accumulation in node / out-edge order
        }
      }
    }
    std::vector<uint32_t> idx(keys.size());
    std::iota(idx.begin(), idx.end(), 0u);
    std::sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return keys[x] < keys[y]; });
    ma.resize(keys.size());
    mb.resize(keys.size());
    wsum.resize(keys.size());
    for (size_t i = 0; i < idx.size(); ++i) {
      ma[i] = (uint32_t)(keys[idx[i]] / T);
      mb[i] = (uint32_t)(keys[idx[i]] % T);
      wsum[i] = sums[idx[i]];
    }
  }
  uint32_t n_cc = 0;
  const std::vector<uint32_t> cc = connected_components(T, ma, mb, nullptr, &n_cc);
  S.n_meta_components = n_cc;
  std::vector<uint64_t> cc_nodes(n_cc, 0);
  for (uint32_t t = 0; t < T; ++t) cc_nodes[cc[t]] += nodes_in_track[t];
  std::vector<uint32_t> gc(cc);
  uint32_t next_label = n_cc;
  const uint32_t max_nodes = S.n_images_seen;
  std::vector<std::vector<MetaEdge>> per_cc;
  std::vector<int32_t> big_slot(n_cc, -1);
  for (uint32_t c = 0; c < n_cc; ++c)
    if (cc_nodes[c] > max_nodes) {
      big_slot[c] = (int32_t)per_cc.size();
      per_cc.emplace_back();
    }
  S.n_oversized_meta_components = (uint32_t)per_cc.size();
  if (!per_cc.empty()) {
    for (size_t i = 0; i < ma.size(); ++i) {
      if (!(ma[i] < mb[i])) continue;  // undirected list, weight int(100 * sum sim) (solve.cc:327-330)
      const int32_t s = big_slot[cc[ma[i]]];
      if (s < 0) continue;
      per_cc[s].push_back(MetaEdge{ma[i], mb[i], (int64_t)(100.0 * wsum[i])});
    }
    for (auto& edges : per_cc) {
      std::vector<std::vector<uint32_t>> groups;
      recursive_cut(edges, nodes_in_track, max_nodes, &groups);
      for (const auto& g : groups) {
        for (uint32_t t : g) gc[t] = next_label;
        ++next_label;
      }
      S.n_cut_groups += (uint32_t)groups.size();
    }
  }
  std::vector<uint8_t> keep(ma.size());
  for (size_t i = 0; i < ma.size(); ++i) keep[i] = gc[ma[i]] == gc[mb[i]];
  uint32_t n_final = 0;
  const std::vector<uint32_t> final_cc = connected_components(T, ma, mb, &keep, &n_final);
  hs->comp.resize(N);
  for (uint32_t v = 0; v < N; ++v) hs->comp[v] = final_cc[hs->track[v]];
  hs->C = n_final;
  S.graph_cut_ms = ms_since(t_cut);
  // ---- H5: dispatch list (solve.cc:594-604) ----------------------------------------------------
  std::vector<uint32_t> sizes(n_final, 0);
  for (uint32_t v = 0; v < N; ++v) ++sizes[hs->comp[v]];
  hs->comp_order.resize(n_final);
  std::iota(hs->comp_order.begin(), hs->comp_order.end(), 0u);
  std::sort(hs->comp_order.begin(), hs->comp_order.end(), [&](uint32_t x, uint32_t y) {  // sort + reverse on (size, idx)
    if (sizes[x] != sizes[y]) return sizes[x] > sizes[y];
    return x > y;
  });
  std::vector<uint32_t> slot_of(n_final);
  for (uint32_t s = 0; s < n_final; ++s) slot_of[hs->comp_order[s]] = s;
  hs->comp_ptr.assign((size_t)n_final + 1, 0);
  for (uint32_t s = 0; s < n_final; ++s) hs->comp_ptr[s + 1] = hs->comp_ptr[s] + sizes[hs->comp_order[s]];
  hs->comp_nodes.resize(N);
  {
    std::vector<uint32_t> fill(hs->comp_ptr.begin(), hs->comp_ptr.end() - 1);
    for (uint32_t v = 0; v < N; ++v) hs->comp_nodes[fill[slot_of[hs->comp[v]]]++] = v;  // ascending node index
  }
  S.n_nodes = N;
  S.n_edges = hs->E;
  S.n_tracks = T;
  S.n_components = n_final;
  S.max_component_size = n_final ? sizes[hs->comp_order[0]] : 0;
  *out = hs;
  if (sz) *sz = S;
  return LFR_OK;
}

int lfr_host_stage_export(const lfr_host_stage* hs, uint32_t* row_ptr, lfr_edge* edges, uint32_t* track, uint32_t* comp,
                          uint8_t* is_root, uint32_t* comp_ptr, uint32_t* comp_nodes, uint32_t* comp_order,
                          uint32_t* node_image, uint32_t* node_feat) {
  if (!hs) return LFR_EINVAL;
  auto cp = [](void* dst, const void* src, size_t bytes) {
    if (dst && bytes) std::memcpy(dst, src, bytes);
  };
  cp(row_ptr, hs->row_ptr.data(), hs->row_ptr.size() * 4);
  cp(edges, hs->edges.data(), hs->edges.size() * sizeof(lfr_edge));
  cp(track, hs->track.data(), hs->track.size() * 4);
  cp(comp, hs->comp.data(), hs->comp.size() * 4);
  cp(is_root, hs->is_root.data(), hs->is_root.size());
  cp(comp_ptr, hs->comp_ptr.data(), hs->comp_ptr.size() * 4);
  cp(comp_nodes, hs->comp_nodes.data(), hs->comp_nodes.size() * 4);
  cp(comp_order, hs->comp_order.data(), hs->comp_order.size() * 4);
  cp(node_image, hs->node_image.data(), hs->node_image.size() * 4);
  cp(node_feat, hs->node_feat.data(), hs->node_feat.size() * 4);
  return LFR_OK;
}

void lfr_host_stage_destroy(lfr_host_stage* hs) { delete hs; }

}  // extern "C"
