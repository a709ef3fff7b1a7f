"""Consumer side of the SolutionFile: apply the refined
displacements to keypoints the way the reference's reconstruction scripts do,
without the per-displacement Python loop.

reconstruction-scripts/colmap_utils.py:104-137 walks every Displacement message
of an image and writes `displacements[feature_idx] = [dj, di]`, then

    keypoints[:, :2] += displacements * fact * 16      # 1 solver unit = 16 px
    keypoints[:, :2] += 0.5

(the multiplication by `fact` happens in float32 before the multiplication by
16, and keypoints are float32; both are reproduced so the resulting blob is
bit-identical).  Later displacements for the same feature overwrite earlier
ones, as in the reference's loop.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import wire


def complete_keypoints(keypoints: np.ndarray) -> np.ndarray:
    """colmap_utils.py:64-74 — pad (x, y[, scale]) to COLMAP's 4 columns."""
    k = np.asarray(keypoints)
    if k.shape[1] == 2:
        return np.hstack([k, np.ones([k.shape[0], 1]), np.zeros([k.shape[0], 1])])
    if k.shape[1] == 3:
        return np.hstack([k, np.zeros([k.shape[0], 1])])
    return k


def apply_to_keypoints(keypoints: np.ndarray, feature_idx: Optional[np.ndarray], di: Optional[np.ndarray],
                       dj: Optional[np.ndarray], fact: float = 1.0) -> np.ndarray:
    """One image: returns the float32 [n, 4] keypoint array COLMAP's database gets
    (colmap_utils.py:118-137).  `feature_idx/di/dj` = the image's Displacement
    list, or None when the image has no entry in the SolutionFile."""
    kp = complete_keypoints(keypoints[:, :3] if keypoints.shape[0] else np.zeros([0, 4])).astype(np.float32)
    n = kp.shape[0]
    if feature_idx is not None:
        disp = np.zeros([n, 2], dtype=np.float32)
        fi = np.asarray(feature_idx, dtype=np.int64)
        # last write wins, like the reference's sequential assignments
        disp[fi, 0] = np.asarray(dj, dtype=np.float32)
        disp[fi, 1] = np.asarray(di, dtype=np.float32)
        disp *= np.float32(fact)
        kp[:, :2] += disp * 16
    kp[:, :2] += 0.5
    return kp


def apply_solution(keypoints_by_image: Dict[str, np.ndarray], solution_bytes: bytes) -> Dict[str, np.ndarray]:
    """All images: {image_name: keypoints [n, >=2]} + SolutionFile bytes ->
    {image_name: float32 [n, 4]} ready for `array_to_blob` (colmap_utils.py:143)."""
    sol = {name: (fact, fi, di, dj) for name, fact, fi, di, dj in wire.decode_solution(solution_bytes)}
    out = {}
    for name, kp in keypoints_by_image.items():
        if name in sol:
            fact, fi, di, dj = sol[name]
            out[name] = apply_to_keypoints(kp, fi, di, dj, fact)
        else:
            out[name] = apply_to_keypoints(kp, np.zeros(0, np.int64), np.zeros(0), np.zeros(0), 1.0)
    return out
