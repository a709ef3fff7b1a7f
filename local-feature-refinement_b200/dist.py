"""Multi-GPU, one PROCESS per GPU: independent components partitioned across ranks.

(The single-process route — one call driving several devices, no collective at all — is
lfr_solve_multi() in the C ABI, include/lfr.h; `solve --gpus N` uses that.  This module is the
torchrun / torch.distributed variant for hosts that run one rank per GPU.)

Every component is a self-contained problem (cross-component edges are dropped,
solve.cc:123-125) and none exceeds #images nodes (solve.cc:586), so no component
ever spans devices: ranks need no collective inside the LM loop.  Each rank extracts the
sub-graph of its components (the 80-byte edge records — the bulk — are partitioned, never
replicated), solves it with lfr_solve on its own device, and the results are
combined by ONE all-reduce(sum) of one packed buffer [positions | per-component stats | totals]
(ranks write disjoint entries of a zero-initialised buffer, so the sum is exact and the result is
bitwise identical to a 1-GPU solve).
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np

from .graph import EDGE_DTYPE, MatchGraph, Problem


def lpt_partition(weights: np.ndarray, n_parts: int) -> List[np.ndarray]:
    """Longest-processing-time-first bin packing of dispatch slots; mirrors the
    reference's largest-first queue (solve.cc:599-604).  Deterministic."""
    order = np.argsort(-weights, kind="stable")
    load = [0] * n_parts
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for s in order.tolist():
        k = min(range(n_parts), key=lambda i: (load[i], i))
        parts[k].append(s)
        load[k] += int(weights[s])
    return [np.array(sorted(x), dtype=np.int64) for x in parts]


def slot_weights(p: Problem) -> np.ndarray:
    """Work estimate of a dispatch slot = its directed edges (0 for size-1 components)."""
    deg = p.graph.row_ptr[1:].astype(np.int64) - p.graph.row_ptr[:-1].astype(np.int64)
    node_w = deg[p.comp_nodes.astype(np.int64)]
    csum = np.concatenate([[0], np.cumsum(node_w)])
    ptr = p.comp_ptr.astype(np.int64)
    w = csum[ptr[1:]] - csum[ptr[:-1]]
    w[(ptr[1:] - ptr[:-1]) <= 1] = 0
    return w


def shard_problem(p: Problem, slots: np.ndarray) -> Tuple[Problem, np.ndarray]:
    """Sub-problem holding only the given dispatch slots, nodes re-indexed
    locally.  Returns (sub-problem, global node index of each local node)."""
    g = p.graph
    ptr = p.comp_ptr.astype(np.int64)
    sizes = (ptr[1:] - ptr[:-1])[slots]
    if slots.shape[0]:
        node_lists = [p.comp_nodes[ptr[s]:ptr[s + 1]].astype(np.int64) for s in slots.tolist()]
        gnodes = np.concatenate(node_lists) if node_lists else np.zeros(0, np.int64)
    else:
        gnodes = np.zeros(0, np.int64)
    n_loc = gnodes.shape[0]
    local_of = np.full(g.n_nodes, -1, dtype=np.int64)
    local_of[gnodes] = np.arange(n_loc)
    rp = g.row_ptr.astype(np.int64)
    deg = rp[gnodes + 1] - rp[gnodes]
    new_rp = np.zeros(n_loc + 1, dtype=np.int64)
    np.cumsum(deg, out=new_rp[1:])
    # gather the out-edge rows of the local nodes
    idx = np.repeat(rp[gnodes] - new_rp[:-1], deg) + np.arange(int(new_rp[-1]))
    e = g.edges[idx].copy()
    dst_loc = local_of[e["dst"].astype(np.int64)]
    # edges leaving the shard are cross-component by construction (skipped by the
    # solve): point them at their own source with a foreign component id
    src_loc = np.repeat(np.arange(n_loc), deg)
    outside = dst_loc < 0
    keep = ~outside
    e = e[keep]
    e["dst"] = dst_loc[keep].astype(np.uint32)
    kept_deg = np.bincount(src_loc[keep], minlength=n_loc)
    new_rp = np.zeros(n_loc + 1, dtype=np.uint32)
    np.cumsum(kept_deg, out=new_rp[1:])
    sub_g = MatchGraph(
        n_nodes=n_loc, node_image=g.node_image[gnodes], node_feat=g.node_feat[gnodes],
        und_sim=np.zeros(0), und_n1=np.zeros(0, np.int64), und_n2=np.zeros(0, np.int64),
        row_ptr=new_rp, edges=np.ascontiguousarray(e, dtype=EDGE_DTYPE), image_names=g.image_names,
        image_fact=g.image_fact, n_images=g.n_images)
    comp_ptr = np.zeros(slots.shape[0] + 1, dtype=np.uint32)
    np.cumsum(sizes, out=comp_ptr[1:])
    sub = Problem(graph=sub_g, track=p.track[gnodes], comp=p.comp[gnodes], is_root=p.is_root[gnodes],
                  comp_ptr=comp_ptr, comp_nodes=np.arange(n_loc, dtype=np.uint32),
                  comp_order=p.comp_order[slots] if p.comp_order.shape[0] else p.comp_order, info=dict(p.info))
    return sub, gnodes


def solve_sharded(p: Problem, rank: int, world: int, solve_fn: Callable, all_reduce_sum: Callable):
    """Core of the N>1 path, independent of the process-group backend:
    partition -> local solve -> all-reduce.  `solve_fn(sub_problem)` returns
    (positions, stats); `all_reduce_sum(np.ndarray)` sums in place across ranks."""
    parts = lpt_partition(slot_weights(p), world)
    sub, gnodes = shard_problem(p, parts[rank])
    pos_loc, st = solve_fn(sub)
    N, C = p.graph.n_nodes, p.n_components
    # one packed buffer -> one collective: positions [2N] | iterations, termination, cost0, cost1 [C each] | 4 totals
    buf = np.zeros(2 * N + 4 * C + 4, dtype=np.float64)
    pos = buf[:2 * N].reshape(N, 2)
    pos[gnodes] = pos_loc
    mine = parts[rank]
    o = 2 * N
    buf[o + mine] = st["iterations"]                  # small integers: exact in float64
    buf[o + C + mine] = st["termination"]
    buf[o + 2 * C + mine] = st["initial_cost"]
    buf[o + 3 * C + mine] = st["final_cost"]
    buf[o + 4 * C:] = [float(st["total_iterations"]), float(st["total_line_search_steps"]),
                       float(st["n_solved"]), float(st.get("n_kernel_launches", 0))]
    all_reduce_sum(buf)
    scal = buf[o + 4 * C:]
    out = dict(iterations=buf[o:o + C].astype(np.int32), termination=buf[o + C:o + 2 * C].astype(np.int32),
               initial_cost=buf[o + 2 * C:o + 3 * C].copy(), final_cost=buf[o + 3 * C:o + 4 * C].copy(),
               total_iterations=int(scal[0]), total_line_search_steps=int(scal[1]),
               n_solved=int(scal[2]), n_kernel_launches=int(scal[3]),
               kernel_ms=st.get("kernel_ms", 0.0), h2d_ms=st.get("h2d_ms", 0.0), d2h_ms=st.get("d2h_ms", 0.0),
               total_ms=st.get("total_ms", 0.0), shard_slots=[int(x.shape[0]) for x in parts])
    return pos.copy(), out


def solve_distributed(p: Problem, options=None):
    """Called on every rank of a torchrun launch (cli.py --gpus N)."""
    import os
    import torch
    import torch.distributed as dist
    from .capi import load_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if not dist.is_initialized():
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_b200()
    opts = options if options is not None else lib.default_options()
    opts.device = local

    def all_reduce_sum(a: np.ndarray):
        t = torch.from_numpy(a).cuda(local)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        a[...] = t.cpu().numpy()

    return solve_sharded(p, rank, world, lambda q: lib.solve(q, opts), all_reduce_sum)
