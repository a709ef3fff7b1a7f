"""Producer side of the MatchingFile: what two-view-refinement/
compute_match_graph.py:163-205 does with python-protobuf objects — one `add()` and seven
attribute assignments per grid sample, 18 samples per match — as array appends and one native
encode per part.

    w = MatchGraphWriter(args.output_file)                      # dump_interval = 5000, :77-78
    for pair_idx, ... in enumerate(match list):                 # :95
        ...
        w.add_pair(image_name1, fact1, image_name2, fact2, matches, sim,
                   grid_displacements12, grid_displacements21)  # replaces :163-187
    w.close()                                                   # replaces :189-205

writes byte-identical files (`output_file`, or `output_file.part.N` every `dump_interval` pairs,
with the reference's exact part numbering incl. the trailing part) and, with packed=True, the
flat-array form `output_file.npz` that `solve --matches_file X.npz` reads without parsing.
Field conventions (types.proto:15-21): disp1 = grid_displacements21, disp2 = grid_displacements12,
grid order 3*i + j, (di, dj) = last axis (compute_match_graph.py:175-187).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from . import wire
from .matchset import MatchSet


class MatchGraphWriter:
    def __init__(self, output_file: str, dump_interval: int = 5000, packed: bool = False):
        self.output_file = output_file
        self.dump_interval = int(dump_interval)
        self.packed = packed
        self.part_idx = -1                       # compute_match_graph.py:93
        self.pair_idx = -1
        self._all: List[MatchSet] = []           # every part, for the packed form
        self._reset()

    def _reset(self):
        self._names: List[str] = []
        self._index = {}
        self._img1: List[int] = []
        self._img2: List[int] = []
        self._fact1: List[float] = []
        self._fact2: List[float] = []
        self._ptr = [0]
        self._f1, self._f2, self._sim, self._d1, self._d2 = [], [], [], [], []

    def _intern(self, name: str) -> int:
        i = self._index.get(name)
        if i is None:
            i = self._index[name] = len(self._names)
            self._names.append(name)
        return i

    def add_pair(self, image_name1: str, fact1: float, image_name2: str, fact2: float, matches: np.ndarray,
                 sim: Optional[np.ndarray], grid_displacements12: Optional[np.ndarray],
                 grid_displacements21: Optional[np.ndarray]) -> None:
        """One iteration of the reference's loop body after matching/refinement
        (compute_match_graph.py:163-194).  `matches` [m, 2]; `sim` [m]; grids [m, 3, 3, 2]."""
        self.pair_idx += 1
        m = int(np.asarray(matches).shape[0])
        self._img1.append(self._intern(image_name1))
        self._img2.append(self._intern(image_name2))
        self._fact1.append(fact1)
        self._fact2.append(fact2)
        if m:
            mt = np.asarray(matches)
            self._f1.append(mt[:, 0].astype(np.uint32))
            self._f2.append(mt[:, 1].astype(np.uint32))
            self._sim.append(np.asarray(sim)[:m].astype(np.float32))
            self._d1.append(np.asarray(grid_displacements21).reshape(m, 18).astype(np.float32))   # disp1 <- 21
            self._d2.append(np.asarray(grid_displacements12).reshape(m, 18).astype(np.float32))   # disp2 <- 12
        self._ptr.append(self._ptr[-1] + m)
        if self.pair_idx % self.dump_interval == self.dump_interval - 1:                           # :189-194
            self.part_idx += 1
            self._write("%s.part.%d" % (self.output_file, self.part_idx))
            self._reset()

    def _matchset(self) -> MatchSet:
        cat = lambda xs, dt, shape: (np.concatenate(xs).astype(dt) if xs else np.zeros(shape, dtype=dt))
        return MatchSet(image_names=list(self._names), pair_img1=np.array(self._img1, dtype=np.int64),
                        pair_img2=np.array(self._img2, dtype=np.int64),
                        pair_fact1=np.array(self._fact1, dtype=np.float32), pair_fact2=np.array(self._fact2, dtype=np.float32),
                        pair_ptr=np.array(self._ptr, dtype=np.int64), feat1=cat(self._f1, np.uint32, (0,)),
                        feat2=cat(self._f2, np.uint32, (0,)), sim=cat(self._sim, np.float32, (0,)),
                        disp1=cat(self._d1, np.float32, (0, 18)), disp2=cat(self._d2, np.float32, (0, 18)))

    def _write(self, path: str) -> None:
        ms = self._matchset()
        with open(path, "wb") as fh:
            fh.write(wire.encode_matching_file(ms))
        self._all.append(ms)

    def close(self) -> None:
        """compute_match_graph.py:196-205: a single file when no part was dumped, else one more part."""
        if self.part_idx == -1:
            self._write(self.output_file)
        else:
            self.part_idx += 1
            self._write("%s.part.%d" % (self.output_file, self.part_idx))
        if self.packed:
            MatchSet.concatenate(self._all).save_npz(self.output_file + ".npz")
