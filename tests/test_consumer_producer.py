"""consumer-side apply and the packed producer format."""
import numpy as np

from lfr_b200 import MatchSet, synth, wire
from lfr_b200.consumer import apply_solution, apply_to_keypoints


def reference_apply(keypoints, displacements, fact):
    """reconstruction-scripts/colmap_utils.py:118-137, loop and all."""
    kp = keypoints[:, :3]
    kp = np.hstack([kp, np.zeros([kp.shape[0], 1])]).astype(np.float32)
    disp = np.zeros([kp.shape[0], 2]).astype(np.float32)
    for (feature_idx, di, dj) in displacements:
        disp[feature_idx, :] = [dj, di]
    disp *= fact
    kp[:, :2] += disp * 16
    kp[:, :2] += 0.5
    return kp


def test_apply_matches_reference_loop_bitwise():
    rng = np.random.default_rng(0)
    kp = rng.uniform(0, 1000, size=(50, 3)).astype(np.float64)
    fi = rng.permutation(50)[:30]
    fi = np.concatenate([fi, fi[:3]])                      # repeated features: last write wins
    di = rng.uniform(-1, 1, size=fi.shape[0]).astype(np.float32)
    dj = rng.uniform(-1, 1, size=fi.shape[0]).astype(np.float32)
    fact = np.float32(1.7)
    want = reference_apply(kp, list(zip(fi.tolist(), di.tolist(), dj.tolist())), fact)
    got = apply_to_keypoints(kp, fi, di, dj, fact)
    assert got.dtype == np.float32 and got.tobytes() == want.tobytes()
    # an image absent from the solution only gets the +0.5 shift
    none = apply_to_keypoints(kp, None, None, None)
    assert np.array_equal(none[:, :2], (kp[:, :2].astype(np.float32) + np.float32(0.5)))


def test_apply_solution_from_bytes():
    data = wire.encode_solution(["a.png"], np.array([2.0], np.float32), np.array([0, 2]), np.array([1, 0], np.uint32),
                                np.array([0.25, -0.5], np.float32), np.array([0.125, 0.0], np.float32))
    kps = {"a.png": np.array([[10.0, 20.0], [30.0, 40.0]]), "b.png": np.array([[1.0, 2.0, 3.0]])}
    out = apply_solution(kps, data)
    np.testing.assert_allclose(out["a.png"][0, :2], [10 + 0.0 * 2 * 16 + 0.5, 20 - 0.5 * 2 * 16 + 0.5])
    np.testing.assert_allclose(out["a.png"][1, :2], [30 + 0.125 * 2 * 16 + 0.5, 40 + 0.25 * 2 * 16 + 0.5])
    assert out["a.png"].shape == (2, 4) and out["a.png"][0, 2] == 1.0       # 2-column keypoints get scale 1
    np.testing.assert_allclose(out["b.png"], [[1.5, 2.5, 3.0, 0.0]])


def test_packed_matchset_roundtrip(tmp_path):
    ms = synth.generate("cfg1", scale=0.3)
    path = str(tmp_path / "m.npz")
    ms.save_npz(path)
    back = MatchSet.load_npz(path)
    assert back.image_names == ms.image_names
    for k in ("pair_img1", "pair_img2", "pair_fact1", "pair_fact2", "pair_ptr", "feat1", "feat2", "sim", "disp1", "disp2"):
        assert np.array_equal(getattr(back, k), getattr(ms, k)), k
    assert wire.encode_matching_file(back) == wire.encode_matching_file(ms)
