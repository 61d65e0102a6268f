"""Worker for the world_size-2 gloo test: emulated tensor-parallel forward vs the full oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_pkg  # noqa: E402
from oracle.binding import Oracle  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pkg = load_pkg()
    path = sys.argv[1]
    views = pkg.checkpoint.load_views(path)
    plan = pkg.tp.shard_plan(views, rank, world)
    orc = Oracle()

    def allreduce(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        dist.all_reduce(t)
        return t.numpy()

    def allgather(a):
        t = torch.from_numpy(np.ascontiguousarray(a))
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return torch.cat(outs).numpy()

    S = 16
    toks = [3, 77, 200, 9]
    worst = 0.0
    # both ways of forming the residual stream: x + allreduce(p) (per-op + NCCL path) and the fused device path's
    # rank-order sum over the gathered partials
    for rank_order in (False, True):
        er = pkg.tp.EmulatedRank(views, plan, orc, allreduce, allgather, S, rank_order=rank_order)
        with orc.open(path, S) as om:
            for pos, t in enumerate(toks):
                full = om.forward(t, pos)
                mine = er.forward(t, pos)
                worst = max(worst, float(np.abs(full - mine).max()))
                assert int(full.argmax()) == int(mine.argmax()), (pos, rank)
                assert np.all(np.abs(full - mine) <= 1e-2 + 1e-3 * np.abs(full)), (pos, rank, worst)
                if rank_order:  # every rank must hold the same bits
                    tt = torch.from_numpy(mine.copy())
                    ref = tt.clone()
                    dist.broadcast(ref, src=0)
                    assert torch.equal(tt, ref), (pos, rank)
    # every rank sees identical logits (all-reduce / all-gather results are replicated)
    t = torch.tensor([worst], dtype=torch.float64)
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert lo.item() == hi.item()
    if rank == 0:
        print(f"TP_OK world={world} worst={worst:.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
