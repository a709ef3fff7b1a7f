"""Host graph stage (graph.py) against the literal restatement of solve.cc's
containers and loops (oracle/host_stage_ref.py) on small random inputs."""
import importlib.util
import os

import numpy as np
import pytest

from lfr_b200 import build_graph, build_problem, compute_tracks, select_roots, separate_meta_graph, synth
from lfr_b200.graph import edge_sources

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("host_stage_ref", os.path.join(ROOT, "oracle", "host_stage_ref.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def as_pairs(ms):
    out = []
    for p in range(ms.n_pairs):
        ms_ = [(int(ms.feat1[m]), int(ms.feat2[m]), float(ms.sim[m]), ms.disp1[m].tolist(), ms.disp2[m].tolist())
               for m in range(int(ms.pair_ptr[p]), int(ms.pair_ptr[p + 1]))]
        out.append((ms.image_names[int(ms.pair_img1[p])], ms.image_names[int(ms.pair_img2[p])], ms_))
    return out


CASES = [("cfg1", 1.0, None), ("cfg1", 0.5, 5), ("cfg2", 0.02, 9), ("cfg4", 0.01, 3)]


@pytest.mark.parametrize("cfg,scale,seed", CASES)
def test_stage_matches_literal_restatement(cfg, scale, seed):
    ms = synth.generate(cfg, scale=scale, seed=seed)
    # similarity ties exercise the (sim, n1, n2) tie-breaks
    ms.sim[:] = np.round(ms.sim * 20) / 20
    nodes, edges, images = ref.build_graph(as_pairs(ms))
    g = build_graph(ms)                      # the numpy stage, function by function
    assert g.n_nodes == len(nodes) and g.n_images == len(images)
    assert [ms.image_names[i] for i in g.node_image.tolist()] == [n["image_name"] for n in nodes]
    assert g.node_feat.tolist() == [n["feature_idx"] for n in nodes]
    src = edge_sources(g)
    for v in range(0, g.n_nodes, max(1, g.n_nodes // 50)):          # out-edge order and payload
        mine = np.nonzero(src == v)[0]
        assert g.edges["dst"][mine].tolist() == [e[0] for e in nodes[v]["out_edges"]]
        for k, e in zip(mine.tolist(), nodes[v]["out_edges"]):
            assert np.array_equal(g.edges["flow"][k], np.array(e[2], np.float32)) and float(g.edges["sim"][k]) == e[1]
    track_ref, n_tracks = ref.tracks(nodes, edges)
    track = compute_tracks(g)
    assert track.tolist() == track_ref
    assert select_roots(g, track).astype(bool).tolist() == ref.roots(nodes, track_ref, n_tracks)
    # with an unreachable size cap no cut happens: components = meta-graph connected components
    comp = separate_meta_graph(g, track, 10 ** 9)
    cc_ref, _, _ = ref.meta_components(nodes, track_ref, n_tracks)
    assert comp.tolist() == [cc_ref[t] for t in track_ref]


@pytest.mark.parametrize("cfg,scale", [("cfg1", 1.0), ("cfg2", 0.05), ("cfg4", 0.02)])
def test_partition_invariants(cfg, scale):
    ms = synth.generate(cfg, scale=scale)
    p = build_problem(ms)
    g = p.graph
    sizes = np.bincount(p.comp.astype(np.int64))
    assert sizes.max() <= g.n_images                                  # solve.cc:586 cap
    # tracks are never split; one node per image inside a track (solve.cc:507-511)
    assert np.all(np.bincount(p.track.astype(np.int64) * 0 + 0) >= 0)
    tc = {}
    for t, c in zip(p.track.tolist(), p.comp.tolist()):
        assert tc.setdefault(t, c) == c
    key = p.track.astype(np.int64) * (int(g.node_image.max()) + 1) + g.node_image
    assert np.unique(key).shape[0] == g.n_nodes
    # exactly one root per track
    assert np.array_equal(np.bincount(p.track.astype(np.int64), weights=p.is_root), np.ones(int(p.track.max()) + 1))
    # dispatch list: sort+reverse on (size, idx); nodes ascending inside a component
    d_ref = ref.dispatch(p.comp.tolist())
    assert [c for c, _ in d_ref] == p.comp_order.tolist()
    for slot, (c, nodes_c) in enumerate(d_ref[:200]):
        assert p.comp_nodes[p.comp_ptr[slot]:p.comp_ptr[slot + 1]].tolist() == nodes_c


@pytest.mark.parametrize("cfg,scale,seed", [("cfg1", 1.0, None), ("cfg2", 0.1, 4), ("cfg4", 0.1, 8), ("ring60", 0.5, 2), ("ring200", 0.3, 5),
                                            ("cfg2", 1.0, None)])
def test_native_host_stage_equals_numpy_stage(cfg, scale, seed):
    """csrc/lfr_host.cc (include/lfr_host.h) against graph.py, array for array."""
    ms = synth.generate(cfg, scale=scale, seed=seed)
    ms.sim[:] = np.round(ms.sim * 50) / 50          # similarity ties
    a = build_problem(ms, native=True)
    b = build_problem(ms, native=False)
    assert a.info["host_stage"] == "native" and b.info["host_stage"] == "numpy"
    for k in ("track", "comp", "is_root", "comp_ptr", "comp_nodes", "comp_order"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    assert np.array_equal(a.graph.row_ptr, b.graph.row_ptr) and a.graph.edges.tobytes() == b.graph.edges.tobytes()
    assert np.array_equal(a.graph.node_image, b.graph.node_image) and np.array_equal(a.graph.node_feat, b.graph.node_feat)
    assert a.graph.n_images == b.graph.n_images and a.graph.image_fact == b.graph.image_fact
    for k in ("n_tracks", "n_components", "max_component_size", "n_meta_components", "n_oversized_meta_components",
              "n_cut_groups"):
        assert a.info[k] == b.info[k], k
    pb = build_problem(ms, banned_images=[ms.image_names[1]], native=True)
    qb = build_problem(ms, banned_images=[ms.image_names[1]], native=False)
    assert np.array_equal(pb.comp, qb.comp) and np.array_equal(pb.comp_nodes, qb.comp_nodes)
    assert pb.graph.n_images == qb.graph.n_images and pb.graph.image_fact == qb.graph.image_fact


def test_banned_images_and_empty_input():
    ms = synth.generate("cfg1")
    p = build_problem(ms, banned_images=["0001.png"])
    assert p.graph.n_images == 2
    names = {ms.image_names[i] for i in p.graph.node_image.tolist()}
    assert "0001.png" not in names
    p0 = build_problem(ms, banned_images=ms.image_names)
    assert p0.graph.n_nodes == 0 and p0.n_components == 0


@pytest.mark.parametrize("cfg,scale", [("cfg2", 1.0), ("cfg4", 0.3), ("cfg5", 0.03)])
def test_native_host_stage_does_not_depend_on_the_thread_count(cfg, scale, monkeypatch):
    """The recursive cut shares its large sub-problems between threads, the edge records are copied by
    several threads: every output array is the same for 1, 3 and 8 of them."""
    ms = synth.generate(cfg, scale=scale)
    ref = None
    for n_thr in ("1", "3", "8"):
        monkeypatch.setenv("LFR_HOST_THREADS", n_thr)
        p = build_problem(ms, native=True)
        got = [p.track, p.comp, p.is_root, p.comp_ptr, p.comp_nodes, p.comp_order, p.graph.row_ptr,
               p.graph.edges.view(np.uint8), np.array([p.info[k] for k in ("n_tracks", "n_components", "n_cut_groups")])]
        if ref is None:
            ref = got
        else:
            for a, b in zip(ref, got):
                assert np.array_equal(a, b)
