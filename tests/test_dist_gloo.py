"""N>1 host path (dist.py) on CPU: two gloo ranks partition the components,
each solves its shard, all-reduce combines.  The per-shard solver here is the
oracle (test injection); on GPUs it is lfr_solve on each rank's device."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import get_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_util import load_oracle
    from lfr_b200 import build_problem, synth
    from lfr_b200.dist import solve_sharded
    orc = load_oracle()
    p = build_problem(synth.generate("cfg1"))

    def all_reduce_sum(a):
        t = torch.from_numpy(a)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)

    pos, st = solve_sharded(p, rank, world, lambda q: orc.solve(q, orc.default_options(n_threads=1)), all_reduce_sum)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), pos=pos, iters=st["iterations"], term=st["termination"],
             total=st["total_iterations"], shard=np.array(st["shard_slots"]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process(tmp_path, oracle):
    world = 2
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    _, p = get_problem("cfg1")
    pos1, st1 = oracle.solve(p, oracle.default_options(n_threads=1))
    r0 = np.load(tmp_path / "rank0.npz")
    r1 = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(r0["pos"], r1["pos"])
    assert np.array_equal(r0["pos"], pos1)                  # bitwise: components never span ranks
    assert np.array_equal(r0["iters"], st1["iterations"]) and np.array_equal(r0["term"], st1["termination"])
    assert int(r0["total"]) == st1["total_iterations"]
    assert r0["shard"].sum() == p.n_components and r0["shard"].min() > 0


def test_lpt_partition_and_shard_extraction(oracle):
    from lfr_b200.dist import lpt_partition, shard_problem, slot_weights
    _, p = get_problem("cfg1")
    w = slot_weights(p)
    parts = lpt_partition(w, 4)
    allslots = np.sort(np.concatenate(parts))
    assert np.array_equal(allslots, np.arange(p.n_components))
    loads = [int(w[x].sum()) for x in parts]
    assert max(loads) - min(loads) <= int(w.max())
    sub, gnodes = shard_problem(p, parts[1])
    assert sub.graph.n_nodes == gnodes.shape[0] and int(sub.comp_ptr[-1]) == gnodes.shape[0]
    pos_full, _ = oracle.solve(p, oracle.default_options(n_threads=1))
    pos_sub, _ = oracle.solve(sub, oracle.default_options(n_threads=1))
    assert np.array_equal(pos_sub, pos_full[gnodes])
