// colmap/base/graph_cut.h — SHIM (test infrastructure written for this repo; not COLMAP).
//
// colmap::ComputeNormalizedMinGraphCut(edges, weights, num_parts) as solve.cc:192 calls it.  COLMAP
// answers it with its bundled Graclus, which exists neither in the reference repository nor in
// this image, so NO build here can reproduce the reference's partition of an oversized
// meta-component.  The shim answers with the one deterministic stand-in this repository defines
// for that call (local-feature-refinement_b200/csrc/lfr_cut.h: a function of the edge list and the
// integer weights only, like the original) — deliberately the same definition the product host
// stage uses, so that everything AROUND the cut (solve.cc:185-373) can be compared exactly.
#ifndef LFR_SHIM_COLMAP_GRAPH_CUT_H_
#define LFR_SHIM_COLMAP_GRAPH_CUT_H_
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../../../local-feature-refinement_b200/csrc/lfr_cut.h"

namespace colmap {

inline std::unordered_map<int, int> ComputeNormalizedMinGraphCut(const std::vector<std::pair<int, int>>& edges,
                                                                 const std::vector<int>& weights, const int num_parts) {
  (void)num_parts;  // the reference only ever asks for 2 (solve.cc:191)
  std::vector<lfr::CutEdge> ce(edges.size());
  for (size_t i = 0; i < edges.size(); ++i)
    ce[i] = lfr::CutEdge{(uint32_t)edges[i].first, (uint32_t)edges[i].second, (int64_t)weights[i]};
  lfr::CutWorkspace W;
  std::unordered_map<int, int> out;
  if (ce.empty()) return out;
  lfr::two_way_cut(ce.data(), ce.size(), W);
  for (size_t i = 0; i < W.nodes.size(); ++i) out[(int)W.nodes[i]] = (int)W.side[i];
  return out;
}

}  // namespace colmap
#endif
