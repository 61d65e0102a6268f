"""2-GPU worker: TP=2 decode (per-op kernels + NCCL) vs the oracle on the same tokens."""
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
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ.update(QWEN_CUDA_DEVICE=str(local), QWEN_CUDA_TP_RANK=str(rank), QWEN_CUDA_TP_SIZE=str(world))
    pkg = load_pkg()
    ql = pkg.QwenLib()
    path = sys.argv[1]
    gm = ql.open(path, 64)
    pkg.tp.init_tensor_parallel(ql, gm, rank, world, dist)
    orc = Oracle()
    toks = np.random.default_rng(2).integers(0, gm.p.vocab_size, size=24)
    with orc.open(path, 64) as om:
        for pos, t in enumerate(toks):
            lg, lo = gm.forward(int(t), pos), om.forward(int(t), pos)
            assert int(lg.argmax()) == int(lo.argmax()), (rank, pos)
            assert np.abs(lg - lo).max() <= 0.05 * lo.std(), (rank, pos, float(np.abs(lg - lo).max()))
    chain = gm.decode_greedy(17, 24, 16)
    t = torch.tensor(chain.tolist(), device="cuda")
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t, ref), "ranks disagree on the greedy chain"
    gm.close()
    dist.barrier()
    if rank == 0:
        print("TP_GPU_OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
