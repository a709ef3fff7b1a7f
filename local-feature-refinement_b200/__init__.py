"""B200-native multi-view local-feature refinement solve.

Drop-in for the hot path of mihaidusmanu/local-feature-refinement
(multi-view-refinement/{cost,graph,solve}.cc): Python host code mirroring
solve.cc's main() around a C-ABI library (include/lfr.h) whose solve is
hand-written sm_100a CUDA.  Import as `lfr_b200` (the directory name carries a
hyphen; the `lfr_b200` shim package at the repo root points here).
"""
from .matchset import MatchSet  # noqa: F401
from .graph import (EDGE_DTYPE, MatchGraph, Problem, build_graph, build_problem,  # noqa: F401
                    compute_tracks, select_roots, separate_meta_graph, refined_track_count)
from . import synth  # noqa: F401

__all__ = ["MatchSet", "MatchGraph", "Problem", "build_graph", "build_problem", "compute_tracks",
           "select_roots", "separate_meta_graph", "refined_track_count", "synth", "EDGE_DTYPE"]
