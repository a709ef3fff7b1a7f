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


def import_features(colmap_path, method_name, database_path, image_path, match_list_path, matches_file, solution_file,
                    run_colmap: bool = True):
    """Drop-in for reconstruction-scripts/colmap_utils.py:77-223 `import_features`: clears the
    feature tables of the COLMAP database, writes every image's (refined) keypoints as float32
    [n, 4] blobs (colmap_utils.py:104-148) and the raw matches of every image pair as uint32 [m, 2]
    blobs (colmap_utils.py:150-190: `.part.N` files, first occurrence of a pair id wins, columns
    swapped when image_id1 > image_id2), then runs `colmap matches_importer` and returns the same
    statistics dict.  The SolutionFile / MatchingFile are decoded by the native codec and applied
    with array operations instead of one Python iteration per displacement / match."""
    import os
    import sqlite3
    import subprocess

    connection = sqlite3.connect(database_path)
    cursor = connection.cursor()
    cursor.execute("SELECT name FROM sqlite_master WHERE type='table' AND name='inlier_matches';")
    inlier_matches_table_exists = cursor.fetchone() is not None
    cursor.execute("DELETE FROM keypoints;")
    cursor.execute("DELETE FROM descriptors;")
    cursor.execute("DELETE FROM matches;")
    cursor.execute("DELETE FROM inlier_matches;" if inlier_matches_table_exists else "DELETE FROM two_view_geometries;")
    connection.commit()
    images = {}
    cursor.execute("SELECT name, image_id FROM images;")
    for row in cursor:
        images[row[0]] = row[1]
    sol = None
    if solution_file is not None:
        with open(solution_file, "rb") as fh:
            # a repeated image name: the reference keeps the LAST message of that name (dict overwrite, :112-114)
            sol = {name: (fact, fi, di, dj) for name, fact, fi, di, dj in wire.decode_solution(fh.read())}
    sum_num_features = 0
    for image_name, image_id in images.items():
        features = np.load(os.path.join(image_path, "%s.%s" % (image_name, method_name)), allow_pickle=True)
        kp_in = features["keypoints"]
        if sol is None:
            kp = complete_keypoints(kp_in[:, :3] if kp_in.shape[0] else np.zeros([0, 4])).astype(np.float32)
            kp[:, :2] += 0.5
        elif image_name in sol:
            fact, fi, di, dj = sol[image_name]
            kp = apply_to_keypoints(kp_in, fi, di, dj, fact)
        else:
            kp = apply_to_keypoints(kp_in, np.zeros(0, np.int64), np.zeros(0), np.zeros(0), 1.0)
        sum_num_features += kp.shape[0]
        assert kp.shape[1] == 4
        cursor.execute("INSERT INTO keypoints(image_id, rows, cols, data) VALUES(?, ?, ?, ?);",
                       (image_id, kp.shape[0], kp.shape[1], kp.tobytes()))
    connection.commit()

    seen = set()
    for path in wire.matches_files(matches_file):
        with open(path, "rb") as fh:
            ms = wire.decode_matching_file(fh.read())
        for p in range(ms.n_pairs):
            id1 = images[ms.image_names[int(ms.pair_img1[p])]]
            id2 = images[ms.image_names[int(ms.pair_img2[p])]]
            pair_id = 2147483647 * id2 + id1 if id1 > id2 else 2147483647 * id1 + id2     # colmap_utils.py:53-57
            if pair_id in seen:
                continue
            seen.add(pair_id)
            a, b = int(ms.pair_ptr[p]), int(ms.pair_ptr[p + 1])
            if b > a:
                cols = (ms.feat2[a:b], ms.feat1[a:b]) if id1 > id2 else (ms.feat1[a:b], ms.feat2[a:b])
                matches = np.stack(cols, axis=1).astype(np.uint32)
            else:
                matches = np.zeros([0, 2])          # the reference's empty float64 array: an empty blob
            cursor.execute("INSERT INTO matches(pair_id, rows, cols, data) VALUES(?, ?, ?, ?);",
                           (pair_id, matches.shape[0], matches.shape[1], matches.tobytes()))
        connection.commit()
    cursor.close()
    connection.close()

    if run_colmap:
        subprocess.call([os.path.join(colmap_path, "colmap"), "matches_importer", "--database_path", database_path,
                         "--match_list_path", match_list_path, "--match_type", "pairs"])
    connection = sqlite3.connect(database_path)
    cursor = connection.cursor()
    cursor.execute("SELECT count(*) FROM images;")
    num_images = next(cursor)[0]
    cursor.execute("SELECT count(*) FROM two_view_geometries WHERE rows > 0;")
    num_inlier_pairs = next(cursor)[0]
    cursor.execute("SELECT sum(rows) FROM two_view_geometries WHERE rows > 0;")
    num_inlier_matches = next(cursor)[0]
    cursor.close()
    connection.close()
    return dict(num_images=num_images, num_inlier_pairs=num_inlier_pairs, num_inlier_matches=num_inlier_matches,
                avg_num_features=(sum_num_features / num_images))
