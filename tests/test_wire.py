"""Wire format (types.proto): the native codec against python-protobuf."""
import os

import numpy as np
import pytest

from lfr_b200 import synth, wire
from proto_runtime import MatchingFile, SolutionFile, matchset_to_proto


@pytest.fixture(scope="module")
def small_ms():
    ms = synth.generate("cfg1", scale=0.25)
    # proto3 corner cases: zero scalars are omitted on the wire
    ms.feat1[0] = 0
    ms.sim[1] = 0.0
    ms.disp1[2, :] = 0.0
    ms.disp2[3, 4] = -0.0
    ms.pair_fact2[0] = 0.0
    return ms


def _assert_same(a, b):
    assert a.image_names == b.image_names
    for k in ("pair_img1", "pair_img2", "pair_fact1", "pair_fact2", "pair_ptr", "feat1", "feat2", "sim", "disp1", "disp2"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k


def test_decode_what_protobuf_writes(small_ms):
    data = matchset_to_proto(small_ms).SerializeToString()
    _assert_same(wire.decode_matching_file(data), small_ms)


def test_encode_is_byte_identical_to_protobuf(small_ms):
    ours = wire.encode_matching_file(small_ms)
    theirs = matchset_to_proto(small_ms).SerializeToString()
    assert ours == theirs
    mf = MatchingFile()
    mf.ParseFromString(ours)
    assert len(mf.image_pairs) == small_ms.n_pairs


def test_short_and_long_displacement_lists():
    """disp lists shorter than 9 leave zeros (solve.cc:460-472); entries beyond the
    9th are ignored (the reference would write out of bounds)."""
    mf = MatchingFile()
    ip = mf.image_pairs.add()
    ip.image_name1, ip.image_name2, ip.fact1, ip.fact2 = "a", "b", 1.0, 2.0
    m = ip.matches.add()
    m.feature_idx1, m.feature_idx2, m.similarity = 7, 9, 0.5
    for g in range(3):
        d = m.disp1.add(); d.di, d.dj = g + 1.0, -(g + 1.0)
    for g in range(11):
        d = m.disp2.add(); d.di, d.dj = 0.25 * g, 0.5 * g
    ms = wire.decode_matching_file(mf.SerializeToString())
    assert ms.n_matches == 1 and ms.image_names == ["a", "b"]
    assert ms.disp1[0, :6].tolist() == [1, -1, 2, -2, 3, -3] and np.all(ms.disp1[0, 6:] == 0)
    assert ms.disp2[0, 16] == 2.0 and ms.disp2[0, 17] == 4.0
    assert float(ms.pair_fact2[0]) == 2.0


def test_unknown_fields_are_skipped_and_garbage_is_rejected(small_ms):
    data = wire.encode_matching_file(small_ms)
    extra = bytes([0x78, 0x05, 0x82, 0x01, 0x03, 1, 2, 3])   # field 15 varint, field 16 bytes
    _assert_same(wire.decode_matching_file(data + extra), small_ms)
    with pytest.raises(wire.ParseError):
        wire.decode_matching_file(data[:-3])
    with pytest.raises(wire.ParseError):
        wire.decode_matching_file(b"\x0a\xff\xff\xff\xff\x0f")
    assert wire.decode_matching_file(b"").n_pairs == 0


def test_displacement_encodings_other_than_the_canonical_one():
    """The decoder takes a shortcut for the usual 10-byte displacement (di then dj, both non-zero); every
    other legal encoding goes through the generic field loop: dj before di, a zero field omitted, an
    unknown field inside, an empty displacement, a repeated field (last value wins)."""
    import struct

    def f32(tag, v):
        return bytes([tag]) + struct.pack("<f", v)

    def lenpref(tag, payload):
        assert len(payload) < 128
        return bytes([tag, len(payload)]) + payload

    disps = [f32(0x0D, 1.5) + f32(0x15, -2.5),            # canonical
             f32(0x15, 4.0) + f32(0x0D, 3.0),             # dj first
             f32(0x0D, 5.0),                              # dj omitted (zero)
             f32(0x15, 6.0),                              # di omitted
             b"",                                         # both zero
             f32(0x0D, 7.0) + bytes([0x18, 0x2A]) + f32(0x15, 8.0),   # unknown varint field 3 in between
             f32(0x0D, 9.0) + f32(0x0D, 10.0) + f32(0x15, 11.0)]     # di twice: the last one counts
    match = bytes([0x08, 3, 0x10, 4]) + f32(0x1D, 0.75) + b"".join(lenpref(0x22, d) for d in disps) \
        + lenpref(0x2A, f32(0x0D, -1.0) + f32(0x15, -3.0))
    pair = lenpref(0x0A, b"x") + f32(0x15, 2.0) + lenpref(0x1A, b"y") + f32(0x25, 4.0) + lenpref(0x2A, match)
    data = lenpref(0x0A, pair)
    ms = wire.decode_matching_file(data)
    assert ms.n_matches == 1 and int(ms.feat1[0]) == 3 and int(ms.feat2[0]) == 4 and float(ms.sim[0]) == 0.75
    assert ms.disp1[0, :14].tolist() == [1.5, -2.5, 3.0, 4.0, 5.0, 0.0, 0.0, 6.0, 0.0, 0.0, 7.0, 8.0, 10.0, 11.0]
    assert ms.disp2[0, :2].tolist() == [-1.0, -3.0] and np.all(ms.disp2[0, 2:] == 0)
    mf = MatchingFile()
    mf.ParseFromString(data)                              # python-protobuf agrees
    got = [(d.di, d.dj) for d in mf.image_pairs[0].matches[0].disp1]
    assert got == [(1.5, -2.5), (3.0, 4.0), (5.0, 0.0), (0.0, 6.0), (0.0, 0.0), (7.0, 8.0), (10.0, 11.0)]


def test_garbage_inside_a_match_is_rejected(small_ms):
    """The pair-level scan skips matches by their length; their contents are checked by the decode."""
    data = bytearray(wire.encode_matching_file(small_ms))
    ms = small_ms
    # corrupt the first displacement of the first match: claim a length that runs past the match
    i = data.index(bytes([0x22, 0x0A, 0x0D]))
    data[i + 1] = 0x7F
    with pytest.raises(wire.ParseError):
        wire.decode_matching_file(bytes(data))
    # a big file (several worker threads): every pair decoded into its own range
    big = synth.generate("cfg2", scale=0.5)
    blob = wire.encode_matching_file(big)
    assert len(blob) > (4 << 20)
    _assert_same(wire.decode_matching_file(blob), big)


def test_part_files(tmp_path, small_ms):
    """`.part.N` every k pairs (compute_match_graph.py:78,189-205), read back like solve.cc:416-424."""
    path = str(tmp_path / "m.pb")
    files = wire.write_matching_file(small_ms, path, pairs_per_part=2)
    assert files == [path + ".part.0", path + ".part.1"] and not os.path.exists(path)
    back = wire.read_matching_file(path)
    _assert_same(back, small_ms)
    # a plain file wins over parts
    wire.write_matching_file(small_ms.select_pairs(np.array([0])), path)
    assert wire.read_matching_file(path).n_pairs == 1


def test_solution_roundtrip_and_bytes():
    names = ["img0.png", "dir/img1.png"]
    fact = np.array([1.0, 0.0], np.float32)          # zero fact omitted
    img_ptr = np.array([0, 3, 5])
    fi = np.array([0, 5, 300, 70000, 1], np.uint32)  # 0 omitted, multi-byte varints
    di = np.array([0.0, 0.5, -0.25, 1e-3, -0.0], np.float32)
    dj = np.array([0.0, -1.0, 0.125, 0.0, 2.0], np.float32)
    data = wire.encode_solution(names, fact, img_ptr, fi, di, dj)
    sf = SolutionFile()
    sf.ParseFromString(data)
    assert [im.image_name for im in sf.images] == names
    assert [d.feature_idx for d in sf.images[0].displacements] == [0, 5, 300]
    assert sf.images[1].displacements[0].feature_idx == 70000
    ref = SolutionFile()
    for i, nm in enumerate(names):
        im = ref.images.add()
        im.image_name, im.fact = nm, float(fact[i])
        for k in range(img_ptr[i], img_ptr[i + 1]):
            d = im.displacements.add()
            d.feature_idx, d.di, d.dj = int(fi[k]), float(di[k]), float(dj[k])
    assert data == ref.SerializeToString()          # e.g. Displacement(0, 0, 0) is "1a 00"
    out = wire.decode_solution(data)
    assert [o[0] for o in out] == names and out[1][1] == 0.0
    assert np.array_equal(out[0][2], fi[:3]) and np.array_equal(out[0][3], di[:3]) and np.array_equal(out[1][4], dj[3:])
