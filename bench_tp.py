"""Tensor-parallel arm of bench.py (N > 1): one process per GPU under torchrun, NCCL all-reduce
after wo / w2 and an all-gather of the logits. Strong scaling: the same
8B-shape decode at context 4096 on 2, 4 or 8 GPUs."""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np


def run(args, base, pkg, shape, shape_name, ctx, K, W, rank, world):
    import torch
    import torch.distributed as dist

    from bench import ClockSampler, cpu_reference_run, ensure_ckpt, log, measured_peak, single_gpu_decode

    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.environ["QWEN_CUDA_DEVICE"] = str(local)
    os.environ["QWEN_CUDA_TP_RANK"] = str(rank)
    os.environ["QWEN_CUDA_TP_SIZE"] = str(world)
    seq_len = ctx + W + K + 48  # room for the CPU arm's thread sweep (reference forward() trusts pos)
    if rank == 0:
        ensure_ckpt(pkg, shape_name)
    dist.barrier()
    path = ensure_ckpt(pkg, shape_name)
    ql = pkg.QwenLib()
    t = time.time()
    gm = ql.open(path, seq_len)
    pkg.tp.init_tensor_parallel(ql, gm, rank, world, dist)
    model_create_s = time.time() - t
    if rank == 0:
        log(f"[bench] tp={world} model_create + nccl init {model_create_s:.1f}s")
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    dev = torch.device("cuda", local)

    def max_over_ranks(x: float) -> float:
        tt = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    dist.barrier()
    torch.cuda.synchronize()
    fused = gm.get_path() == 0
    ms, launches = gm.time_decode(7, ctx, K, W)
    torch.cuda.synchronize()
    dist.barrier()
    ms = max_over_ranks(ms)
    for i in range(W):
        gm.forward_nocopy(7, ctx + i)
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(K):
        if not gm.forward_nocopy(7, ctx + W + i):
            raise SystemExit("forward failed: " + ql.err())
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    # prompt prefill under tensor parallelism (tcgen05 GEMMs per rank + ncclAllReduce of the [T][dim] partial sums): config 5
    # asks for 2048 tokens, the other workloads time 512; best of 2 after one warm-up pass
    prefill = None
    try:
        Tp = 2048 if shape_name == "32b" else 512
        toks = [int(t) for t in np.random.default_rng(0).integers(0, shape.vocab_size, size=Tp)]
        gm.prefill_nocopy(toks, 0)
        best = 1e9
        for _ in range(2):
            dist.barrier()
            t0 = time.perf_counter()
            if not gm.prefill_nocopy(toks, 0):
                raise RuntimeError("prefill failed: " + ql.err())
            best = min(best, max_over_ranks(time.perf_counter() - t0))
        macs = Tp * (shape.weight_elements() - shape.vocab_size * shape.dim)
        prefill = {"workload": f"{shape_name}-prefill{Tp}-tp{world}", "tok_s": Tp / best, "ms": 1e3 * best,
                   "int8_tops_all_gpus": 2 * macs / best / 1e12}
    except Exception as e:
        prefill = {"error": repr(e)}
    gm.close()
    dist.barrier()
    dist.destroy_process_group()  # the other ranks leave now: nothing of theirs may spin beside the one-GPU and CPU legs
    if rank != 0:
        return 0
    if rank == 0:
        # the same workload on ONE GPU, measured now on rank 0's device: the base of the scaling curve
        os.environ.pop("QWEN_CUDA_TP_RANK", None)
        os.environ.pop("QWEN_CUDA_TP_SIZE", None)
        try:
            if args.no_tp1:
                raise RuntimeError("skipped (--no-tp1)")
            if shape.weight_elements() * 17 / 16 + 2 * 4 * shape.n_layers * seq_len * shape.kv_dim > 150e9:
                raise RuntimeError("does not fit one GPU's 180 GB with its KV cache")
            tp1 = single_gpu_decode(pkg, ql, shape_name, ctx, K, W)
        except Exception as e:
            tp1 = {"error": repr(e)}
    if rank == 0:
        tok_s = K / (ms / 1e3)
        bytes_tok = float(np.mean([shape.decode_bytes(ctx + W + i) for i in range(K)]))
        peak, peak_src = measured_peak()
        achieved = bytes_tok / world * tok_s / 1e9  # per-GPU share of the algorithmic bytes
        line = dict(base, value=tok_s, ms_per_step=ms / K, dtype="int8xint8->int32, fp32", clocks=clocks,
                    gpu_launches=launches if fused else launches + K * (2 * shape.n_layers + 1),
                    e2e={"value": K / e2e_s, "unit": "tok/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": shape.vocab_size * 4},
                    roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                              "traffic": None, "peak_source": peak_src, "bytes_per_token": bytes_tok,
                              "kernel": ("k_decode (persistent, 1 launch/token/rank; all-reduce fused into the wo / w2 epilogues as NVLink "
                                         "peer stores) + 1 ncclAllGather of the logits") if fused
                              else "per-op kernels + NCCL all-reduce (peer mapping unavailable)"})
        line["config"]["parallelism"] = f"tp{world}"
        line["config"]["path"] = "mega+peer-stores" if fused else "ops+nccl"
        line["model_create_s"] = model_create_s
        line["prefill"] = prefill
        line["tp1_same_workload"] = tp1
        if "value" in tp1:
            line["vs_tp1_same_workload"] = tok_s / tp1["value"]
        if not args.no_cpu_baseline:
            try:
                r = cpu_reference_run(path, seq_len, ctx, 4, 1, budget_s=40.0)
                line["cpu_baseline"] = {"value": r["value"], "unit": "tok/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                        "all_cores": r["all_cores"], "sweep": r["sweep"]}
            except Exception as e:
                line["cpu_baseline"] = {"value": None, "unit": "tok/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line), flush=True)
    return 0
