"""Slow, literal restatement of solve.cc's host stage — TEST INFRASTRUCTURE ONLY.

Plain Python containers standing in for the reference's std::map / std::set /
std::vector code, statement by statement, so that the vectorised product code
in local-feature-refinement_b200/graph.py can be checked against the
reference's own control flow on small inputs:

    build_graph      solve.cc:438-481   (find_or_create_node :53-65, add_edge graph.cc:19-25)
    tracks           solve.cc:489-549   (union_find_get_root :67-77)
    roots            solve.cc:552-582
    meta_components  solve.cc:258-300   (bfs :162-183)
    dispatch         solve.cc:594-604

The graph-cut (solve.cc:311-343) is not restated: colmap::ComputeNormalizedMinGraphCut
is not part of the reference repository.  Input = list of pairs
(name1, name2, [(feat1, feat2, sim_float32, disp1[18], disp2[18]), ...]).
"""
import numpy as np


def build_graph(pairs, banned=()):
    nodes = {}            # (image_name, feature_idx) -> node_idx      graph_map
    node_list = []        # node_idx -> dict(image_name, feature_idx, out_edges=[(dst, sim, flow)])
    edges = []            # (sim, n1, n2)
    images_set = set()
    for (name1, name2, matches) in pairs:
        if name1 in banned or name2 in banned:
            continue
        images_set.add(name1)
        images_set.add(name2)
        for (f1, f2, sim, disp1, disp2) in matches:
            similarity = float(np.float32(sim))          # double similarity = match.similarity()
            ids = []
            for key in ((name1, int(f1)), (name2, int(f2))):
                if key not in nodes:
                    nodes[key] = len(node_list)
                    node_list.append(dict(image_name=key[0], feature_idx=key[1], out_edges=[]))
                ids.append(nodes[key])
            n1, n2 = ids
            edges.append((similarity, n1, n2))
            node_list[n1]["out_edges"].append((n2, similarity, list(disp2)))   # node1->add_edge(node2, sim, flow_array2)
            node_list[n2]["out_edges"].append((n1, similarity, list(disp1)))   # node2->add_edge(node1, sim, flow_array1)
    return node_list, edges, images_set


def tracks(node_list, edges):
    n_nodes = len(node_list)
    order = sorted(edges)
    order.reverse()
    parent = [-1] * n_nodes
    images_in_track = [{node_list[i]["image_name"]} for i in range(n_nodes)]

    def get_root(i):
        if parent[i] == -1:
            return i
        parent[i] = get_root(parent[i])
        return parent[i]

    for (_, a, b) in order:
        r1, r2 = get_root(a), get_root(b)
        if r1 != r2:
            if images_in_track[r1] & images_in_track[r2]:
                continue
            if len(images_in_track[r1]) < len(images_in_track[r2]):
                parent[r1] = r2
                images_in_track[r2] |= images_in_track[r1]
                images_in_track[r1] = set()
            else:
                parent[r2] = r1
                images_in_track[r1] |= images_in_track[r2]
                images_in_track[r2] = set()
    track_idx = [-1] * n_nodes
    n_tracks = 0
    for i in range(n_nodes):
        if parent[i] == -1:
            track_idx[i] = n_tracks
            n_tracks += 1
    for i in range(n_nodes):
        if track_idx[i] == -1:
            track_idx[i] = track_idx[get_root(i)]
    return track_idx, n_tracks


def roots(node_list, track_idx, n_tracks):
    scores = []
    for i, node in enumerate(node_list):
        score = 0.0
        for (dst, sim, _) in node["out_edges"]:
            if track_idx[i] == track_idx[dst]:
                score += sim
        scores.append((score, i))
    scores.sort()
    scores.reverse()
    is_root = [False] * len(node_list)
    has_root = [False] * n_tracks
    for (_, i) in scores:
        if has_root[track_idx[i]]:
            continue
        is_root[i] = True
        has_root[track_idx[i]] = True
    return is_root


def meta_components(node_list, track_idx, n_tracks):
    """Connected components of the track meta-graph, numbered by first member
    (solve.cc:268-300), and the summed similarities of the meta-edges."""
    meta_edges = [dict() for _ in range(n_tracks)]
    for i, node in enumerate(node_list):
        s = track_idx[i]
        for (dst, sim, _) in node["out_edges"]:
            t = track_idx[dst]
            if s != t:
                meta_edges[s][t] = meta_edges[s].get(t, 0.0) + sim
    comp = [-1] * n_tracks
    n_comp = 0
    for m in range(n_tracks):
        if comp[m] != -1:
            continue
        queue = [m]
        comp[m] = n_comp
        while queue:
            u = queue.pop(0)
            for v in meta_edges[u]:
                if comp[v] == -1:
                    comp[v] = n_comp
                    queue.append(v)
        n_comp += 1
    return comp, n_comp, meta_edges


def dispatch(component_idx):
    n_components = max(component_idx) + 1
    nodes_in_component = [[] for _ in range(n_components)]
    for i, c in enumerate(component_idx):
        nodes_in_component[c].append(i)
    sizes = [(len(nodes_in_component[c]), c) for c in range(n_components)]
    sizes.sort()
    sizes.reverse()
    return [(c, nodes_in_component[c]) for (_, c) in sizes]
