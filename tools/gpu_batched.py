"""Throughput regime: S Fountain-scale scenes solved as one problem (one plan),
to separate the latency bound of a single small scene from the kernel's
throughput when the GPU is full."""
import json, os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import torch
from lfr_b200 import build_problem, refined_track_count, synth
from lfr_b200.capi import Plan, load_b200
from lfr_b200.graph import MatchGraph, Problem

def concat(problems):
    node_off = np.cumsum([0] + [p.graph.n_nodes for p in problems])
    comp_off = np.cumsum([0] + [int(p.comp.max()) + 1 for p in problems])
    track_off = np.cumsum([0] + [int(p.track.max()) + 1 for p in problems])
    edges = np.concatenate([p.graph.edges for p in problems]).copy()
    pos = 0
    for k, p in enumerate(problems):
        e = p.graph.n_edges
        edges["dst"][pos:pos + e] += np.uint32(node_off[k])
        pos += e
    eoff = np.cumsum([0] + [p.graph.n_edges for p in problems])
    row_ptr = np.concatenate([p.graph.row_ptr[:-1].astype(np.int64) + eoff[k] for k, p in enumerate(problems)] + [[eoff[-1]]]).astype(np.uint32)
    g0 = problems[0].graph
    g = MatchGraph(n_nodes=int(node_off[-1]), node_image=np.concatenate([p.graph.node_image for p in problems]),
                   node_feat=np.concatenate([p.graph.node_feat for p in problems]), und_sim=np.zeros(0), und_n1=np.zeros(0, np.int64),
                   und_n2=np.zeros(0, np.int64), row_ptr=row_ptr, edges=edges, image_names=g0.image_names, image_fact=g0.image_fact, n_images=g0.n_images)
    # dispatch list: all slots, largest first (stable merge by size)
    sizes = np.concatenate([np.diff(p.comp_ptr.astype(np.int64)) for p in problems])
    src = np.concatenate([np.full(p.n_components, k) for k, p in enumerate(problems)])
    slot = np.concatenate([np.arange(p.n_components) for p in problems])
    order = np.argsort(-sizes, kind="stable")
    comp_ptr = np.zeros(order.shape[0] + 1, np.uint32); np.cumsum(sizes[order], out=comp_ptr[1:])
    comp_nodes = np.concatenate([problems[src[i]].comp_nodes[problems[src[i]].comp_ptr[slot[i]]:problems[src[i]].comp_ptr[slot[i] + 1]].astype(np.int64) + node_off[src[i]] for i in order]).astype(np.uint32)
    return Problem(graph=g, track=np.concatenate([p.track.astype(np.int64) + track_off[k] for k, p in enumerate(problems)]).astype(np.uint32),
                   comp=np.concatenate([p.comp.astype(np.int64) + comp_off[k] for k, p in enumerate(problems)]).astype(np.uint32),
                   is_root=np.concatenate([p.is_root for p in problems]), comp_ptr=comp_ptr, comp_nodes=comp_nodes,
                   comp_order=np.arange(order.shape[0]), info={})

lib = load_b200()
out = []
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for S in (1, 4, 16, 64):
    ps = [build_problem(synth.generate("cfg2", seed=1002 + 7919 * k if k else None)) for k in range(S)]
    p = concat(ps) if S > 1 else ps[0]
    tracks = sum(refined_track_count(q) for q in ps)
    plan = Plan(lib, p)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        plan.solve(s)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); plan.solve(s); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.mean(ts))
    alg, one = plan.traffic(s)
    pos, st = plan.download(s)
    out.append(dict(scenes=S, components=int(st["n_solved"]), lm_iterations=int(st["total_iterations"]), solve_ms=ms,
                    ms_per_scene=ms / S, tracks_per_s=tracks / (ms / 1e3), algorithmic_GBps=alg / (ms / 1e3) / 1e9,
                    hbm_frac=alg / (ms / 1e3) / 1e9 / 6572.2))
    print(out[-1], flush=True)
    plan.close()
os.makedirs(os.path.join(R, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(R, "gpurun_out", "batched_scenes.json"), "w"), indent=1)
