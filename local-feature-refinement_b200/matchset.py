"""MatchSet — the content of a MatchingFile (types.proto:3-28) as flat arrays.

The reference keeps the parsed protobuf objects around (solve.cc:438-480); here
the same information is a handful of numpy arrays so that the host graph stage
and the packer never loop over matches in Python.

    image_names[i], pair_name1/2 : strings
    pair_img1/2[p]   : image id of each side of pair p (ids index image_names)
    pair_fact1/2[p]  : float32 `fact1` / `fact2` of the pair (types.proto:6,8)
    pair_ptr[p..p+1] : matches of pair p
    feat1/feat2[m]   : uint32 feature_idx1 / feature_idx2 (types.proto:11-12)
    sim[m]           : float32 similarity (types.proto:13)
    disp1/disp2[m]   : float32 [18] = 9 x (di, dj), grid order 3*i+j
                       (types.proto:20-21, compute_match_graph.py:175-187);
                       shorter lists are zero-padded like solve.cc:460-472.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class MatchSet:
    image_names: List[str]
    pair_img1: np.ndarray
    pair_img2: np.ndarray
    pair_fact1: np.ndarray
    pair_fact2: np.ndarray
    pair_ptr: np.ndarray
    feat1: np.ndarray
    feat2: np.ndarray
    sim: np.ndarray
    disp1: np.ndarray
    disp2: np.ndarray
    meta: dict = field(default_factory=dict)

    @property
    def n_pairs(self) -> int:
        return int(self.pair_img1.shape[0])

    @property
    def n_matches(self) -> int:
        return int(self.feat1.shape[0])

    def validate(self) -> None:
        m = self.n_matches
        assert self.pair_ptr.shape == (self.n_pairs + 1,)
        assert int(self.pair_ptr[-1]) == m and int(self.pair_ptr[0]) == 0
        assert self.feat2.shape == (m,) and self.sim.shape == (m,)
        assert self.disp1.shape == (m, 18) and self.disp2.shape == (m, 18)
        assert self.disp1.dtype == np.float32 and self.disp2.dtype == np.float32
        assert self.sim.dtype == np.float32

    # ---- packed on-disk form (the producer can emit this next to,
    # or instead of, the protobuf and skip serialise -> parse) --------------------
    def save_npz(self, path: str) -> None:
        """Write the flat arrays as one .npz (fp32 flows, exactly the wire content)."""
        # names as a fixed-width unicode array: loadable with allow_pickle=False (an object array would
        # make `solve --matches_file X.npz` unpickle whatever file it is handed)
        np.savez(path, image_names=np.array(self.image_names, dtype=np.str_), pair_img1=self.pair_img1,
                 pair_img2=self.pair_img2, pair_fact1=self.pair_fact1, pair_fact2=self.pair_fact2,
                 pair_ptr=self.pair_ptr, feat1=self.feat1, feat2=self.feat2, sim=self.sim, disp1=self.disp1,
                 disp2=self.disp2)

    @staticmethod
    def load_npz(path: str) -> "MatchSet":
        z = np.load(path, allow_pickle=False)
        ms = MatchSet(image_names=[str(x) for x in z["image_names"].tolist()],
                      pair_img1=z["pair_img1"].astype(np.int64), pair_img2=z["pair_img2"].astype(np.int64),
                      pair_fact1=z["pair_fact1"].astype(np.float32), pair_fact2=z["pair_fact2"].astype(np.float32),
                      pair_ptr=z["pair_ptr"].astype(np.int64), feat1=z["feat1"].astype(np.uint32),
                      feat2=z["feat2"].astype(np.uint32), sim=z["sim"].astype(np.float32),
                      disp1=z["disp1"].astype(np.float32), disp2=z["disp2"].astype(np.float32))
        ms.validate()
        return ms

    def without_images(self, banned) -> "MatchSet":
        """Drop every pair touching a banned image (solve.cc:444-446)."""
        banned = set(banned)
        if not banned:
            return self
        keep = np.array([
            (self.image_names[a] not in banned) and (self.image_names[b] not in banned)
            for a, b in zip(self.pair_img1.tolist(), self.pair_img2.tolist())
        ], dtype=bool) if self.n_pairs else np.zeros(0, dtype=bool)
        return self.select_pairs(np.nonzero(keep)[0])

    def select_pairs(self, idx: np.ndarray) -> "MatchSet":
        idx = np.asarray(idx, dtype=np.int64)
        counts = (self.pair_ptr[1:] - self.pair_ptr[:-1])[idx]
        new_ptr = np.zeros(idx.shape[0] + 1, dtype=np.int64)
        np.cumsum(counts, out=new_ptr[1:])
        if idx.shape[0]:
            sel = np.concatenate([np.arange(self.pair_ptr[p], self.pair_ptr[p + 1]) for p in idx.tolist()]) \
                if counts.sum() else np.zeros(0, dtype=np.int64)
        else:
            sel = np.zeros(0, dtype=np.int64)
        sel = sel.astype(np.int64)
        return MatchSet(
            image_names=self.image_names,
            pair_img1=self.pair_img1[idx], pair_img2=self.pair_img2[idx],
            pair_fact1=self.pair_fact1[idx], pair_fact2=self.pair_fact2[idx],
            pair_ptr=new_ptr,
            feat1=self.feat1[sel], feat2=self.feat2[sel], sim=self.sim[sel],
            disp1=self.disp1[sel], disp2=self.disp2[sel], meta=dict(self.meta),
        )

    @staticmethod
    def concatenate(parts: List["MatchSet"]) -> "MatchSet":
        """Concatenate the `.part.N` files of one matching file (solve.cc:416-424):
        image ids are re-interned on the union of names, order of first use."""
        names: List[str] = []
        index = {}
        img1, img2 = [], []
        for ms in parts:
            remap = np.zeros(len(ms.image_names), dtype=np.int64)
            used = np.zeros(len(ms.image_names), dtype=bool)
            used[ms.pair_img1] = True
            used[ms.pair_img2] = True
            # intern in order of first use inside this part
            order = []
            seen = set()
            for a, b in zip(ms.pair_img1.tolist(), ms.pair_img2.tolist()):
                for i in (a, b):
                    if i not in seen:
                        seen.add(i)
                        order.append(i)
            for i in order:
                nm = ms.image_names[i]
                if nm not in index:
                    index[nm] = len(names)
                    names.append(nm)
                remap[i] = index[nm]
            img1.append(remap[ms.pair_img1])
            img2.append(remap[ms.pair_img2])
        ptrs = [np.zeros(1, dtype=np.int64)]
        off = 0
        for ms in parts:
            ptrs.append(ms.pair_ptr[1:].astype(np.int64) + off)
            off += ms.n_matches
        cat = lambda xs, dt: (np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt))
        return MatchSet(
            image_names=names,
            pair_img1=cat(img1, np.int64), pair_img2=cat(img2, np.int64),
            pair_fact1=cat([ms.pair_fact1 for ms in parts], np.float32),
            pair_fact2=cat([ms.pair_fact2 for ms in parts], np.float32),
            pair_ptr=np.concatenate(ptrs),
            feat1=cat([ms.feat1 for ms in parts], np.uint32),
            feat2=cat([ms.feat2 for ms in parts], np.uint32),
            sim=cat([ms.sim for ms in parts], np.float32),
            disp1=(np.concatenate([ms.disp1 for ms in parts]) if parts else np.zeros((0, 18), np.float32)),
            disp2=(np.concatenate([ms.disp2 for ms in parts]) if parts else np.zeros((0, 18), np.float32)),
            meta=dict(parts[0].meta) if parts else {},
        )
