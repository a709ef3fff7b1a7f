"""BASELINE.json configs[3]/[4]: the match graph sharded across N GPUs (one rank
per GPU under torchrun, NCCL).  Components are LPT-packed over the ranks, each
rank solves its shard with lfr_solve on its own device, one all-reduce combines
the positions; rank 0 also solves the whole graph alone and checks that the
sharded result is bitwise identical."""
import json, os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import numpy as np
import torch
import torch.distributed as dist
from lfr_b200 import build_problem, refined_track_count, synth
from lfr_b200.capi import load_b200
from lfr_b200.dist import lpt_partition, shard_problem, slot_weights

name = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_b200()
t0 = time.time()
p = build_problem(synth.generate(name))
t_host = time.time() - t0
parts = lpt_partition(slot_weights(p), world)
sub, gnodes = shard_problem(p, parts[rank])
opts = lib.default_options(device=local)
lib.solve(sub, opts)                        # warm-up (allocations, module load)
dist.barrier(); torch.cuda.synchronize()
t1 = time.perf_counter()
pos_loc, st = lib.solve(sub, opts)          # host buffers -> device -> host
t_solve = time.perf_counter() - t1
pos = torch.zeros((p.graph.n_nodes, 2), dtype=torch.float64, device="cuda")
pos[torch.from_numpy(gnodes).cuda()] = torch.from_numpy(pos_loc).cuda()
t2 = time.perf_counter()
dist.all_reduce(pos, op=dist.ReduceOp.SUM)  # disjoint writes into zeros: exact
torch.cuda.synchronize()
t_ar = time.perf_counter() - t2
tm = torch.tensor([t_solve, t_ar, float(st["kernel_ms"]), float(st["total_iterations"]), float(sub.graph.n_edges)], dtype=torch.float64, device="cuda")
tmax = tm.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
tsum = tm.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
if rank == 0:
    lib.solve(p, opts)
    t3 = time.perf_counter(); pos1, st1 = lib.solve(p, opts); t_single = time.perf_counter() - t3
    same = bool(np.array_equal(pos.cpu().numpy(), pos1))
    tracks = refined_track_count(p)
    out = dict(workload=name, n_gpus=world, nodes=int(p.graph.n_nodes), directed_edges=int(p.graph.n_edges), components=int(p.n_components),
               tracks_refined=tracks, host_stage_s=t_host, shard_edges=[int(x) for x in [tsum[4].item()]],
               solve_ms_max_over_ranks=1e3 * tmax[0].item(), kernel_ms_max_over_ranks=tmax[2].item(), allreduce_ms=1e3 * tmax[1].item(),
               lm_iterations=int(tsum[3].item()), single_gpu_solve_ms=1e3 * t_single, single_gpu_kernel_ms=float(st1["kernel_ms"]),
               tracks_per_s_sharded=tracks / (tmax[0].item() + tmax[1].item()), tracks_per_s_single=tracks / t_single,
               bitwise_identical_to_single_gpu=same)
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(R, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(R, "gpurun_out", "dist_%s_%dgpu.json" % (name, world)), "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
