/*
 * lfr_host.h — C ABI of the native host graph stage ("next" row 2).
 *
 * Everything solve.cc does between the parsed MatchingFile and the solves
 * (solve.cc:438-606): node interning and directed edge lists (:53-65, :474-478),
 * constrained Kruskal -> tracks (:489-549), root selection (:552-582), the track
 * meta-graph, its connected components and the size-capped recursive 2-way cut
 * (:252-373, :185-250) and the dispatch list (:594-604) — from the flat match
 * arrays lfr_wire_decode_matches() produces straight into the arrays of
 * lfr_problem (include/lfr.h).  Host code only; no GPU needed.
 *
 * The 2-way cut stands in for colmap::ComputeNormalizedMinGraphCut (solve.cc:192),
 * which is not part of the reference repository; it is the same deterministic
 * algorithm as local-feature-refinement_b200/graph.py::two_way_cut, and the two
 * implementations are tested to agree exactly.
 */
#ifndef LFR_HOST_H_
#define LFR_HOST_H_

#include <stdint.h>

#include "lfr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lfr_host_stage lfr_host_stage;

typedef struct lfr_host_input {
  uint64_t n_pairs, n_matches;
  uint32_t n_images;            /* image ids are < n_images                      */
  const uint32_t* pair_img1;    /* [n_pairs] image id of side 1 (types.proto:5)   */
  const uint32_t* pair_img2;    /* [n_pairs]                                      */
  const uint8_t* pair_skip;     /* [n_pairs] 1 = pair touches a banned image (solve.cc:444-446), may be NULL */
  const uint64_t* pair_ptr;     /* [n_pairs+1]                                    */
  const uint32_t* feat1;        /* [n_matches]                                    */
  const uint32_t* feat2;
  const float* sim;
  const float* disp1;           /* [n_matches*18]                                 */
  const float* disp2;
  /* optional: caller-owned destination of the 80-byte edge records (the bulk of the output; their
   * number is known up front: 2 x the matches of the non-skipped pairs).  When given and large
   * enough the stage writes them there directly and lfr_host_stage_export() skips its copy. */
  lfr_edge* edges_out;
  uint64_t edges_out_capacity;  /* in records                                     */
} lfr_host_input;

typedef struct lfr_host_sizes {
  uint32_t n_nodes, n_tracks, n_components, n_images_seen, max_track_size, max_component_size;
  uint32_t n_meta_components, n_oversized_meta_components, n_cut_groups, reserved;
  uint64_t n_edges;
  double tracks_ms, graph_cut_ms;   /* solve.cc:487-582 (sort, Kruskal, roots) and :585-589 ("Graph-cut time") */
  double graph_ms, dispatch_ms;     /* node interning + edge lists (:438-481); dispatch list (:594-604)     */
} lfr_host_sizes;

/* Run the whole stage.  Returns 0 or LFR_E*. */
int lfr_host_stage_create(const lfr_host_input* in, lfr_host_stage** out, lfr_host_sizes* sizes);

/* Copy the results into caller-owned arrays (any pointer may be NULL). */
int lfr_host_stage_export(const lfr_host_stage* hs,
                          uint32_t* row_ptr /* [N+1] */, lfr_edge* edges /* [E] */,
                          uint32_t* track /* [N] */, uint32_t* comp /* [N] */, uint8_t* is_root /* [N] */,
                          uint32_t* comp_ptr /* [C+1] */, uint32_t* comp_nodes /* [N] */,
                          uint32_t* comp_order /* [C] component id of each dispatch slot */,
                          uint32_t* node_image /* [N] */, uint32_t* node_feat /* [N] */);

void lfr_host_stage_destroy(lfr_host_stage* hs);

#ifdef __cplusplus
}
#endif
#endif /* LFR_HOST_H_ */
