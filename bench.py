#!/usr/bin/env python
"""bench.py -- decode throughput of the qwen3.c forward hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one forward() = one decoded token. Workloads (BASELINE.json configs):
    4b-decode-ctx4096   Qwen3-4B-shape Q8_0, decode at context 4096, 1 GPU   (headline, default at N=1)
    8b-decode-ctx4096   R1-0528-Qwen3-8B-shape Q8_0, decode at context 4096 (default at N>1)
    1.7b / 0.6b ...     smaller shapes for quick runs
Weights are random-init Q8_0 checkpoints written in the reference's .bin format (no network);
the KV cache is the zero-filled cache the reference also starts from (model.c:360-361): the same
bytes are read whatever their values.

Prints ONE JSON line (see the task contract): `value` = device-timed tok/s with everything resident
in HBM; `e2e` = the same steps through the reference-facing C ABI forward(Model*, token, pos) with
the logits copied back to pinned host memory every step (what the unchanged CLI would see);
`roofline` = algorithmic bytes/token / device time vs the measured HBM peak;
`cpu_baseline` = the reference's own -Ofast/OpenMP build timed on this box's host cores on a
bounded sample. `--impl reference` prints the reference arm's line for the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

KERNEL_TAG = "k_decode_pw"  # the decode kernel single-GPU contexts run (csrc/decode_pw.cuh); ncu captures are tagged with it
CKPT_DIR = os.environ.get("QWEN3_B200_CKPT_DIR", "/tmp/qwen3_b200_ckpt")
WORKLOADS = {
    "4b-decode-ctx4096": ("4b", 4096),
    "8b-decode-ctx4096": ("8b", 4096),
    "32b-decode-ctx32k": ("32b", 32600),   # BASELINE config 5 (TP=8): decode near the reference's 32768-position cap
    "1.7b-decode-ctx512": ("1.7b", 512),
    "0.6b-decode-ctx128": ("0.6b", 128),
    "small-decode-ctx128": ("small", 128),
}


def log(*a):
    print(*a, file=sys.stderr, flush=True)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self._stop_evt = threading.Event()

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        while not self._stop_evt.is_set():
            try:
                out = subprocess.check_output(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                               "--format=csv,noheader,nounits"], text=True, timeout=5)
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 7 for n, v in zip(names, s[3:7]) if v == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ensure_ckpt(pkg, shape_name):
    t = time.time()
    path = pkg.checkpoint.ensure_checkpoint(CKPT_DIR, shape_name, seed=1234, mode="fast")
    log(f"[bench] checkpoint {path} ready in {time.time() - t:.1f}s ({os.path.getsize(path) / 2**30:.2f} GiB)")
    return path


def _omp_set_threads(n):
    """libgomp reads OMP_NUM_THREADS once at load; the sweep sets the team size through its API instead."""
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
        return True
    except OSError:
        return False


def cpu_reference_run(path, seq_len, pos0, steps, warmup, budget_s=150.0):
    """Times the reference's own forward() (oracle/_ref, -Ofast -fopenmp) on the host cores (BASELINE.md
    4.3): sweeps OMP threads over {1, 2, 4, ..., nproc} with one warm + up to two timed calls each (all cores is NOT the
    fastest setting for this code: it opens ~7 k parallel regions per token, reference src/q8.c:24, src/forward.c:107,
    156,174), then times the bounded sample at the BEST thread count. Returns dict(value tok/s at the best setting,
    cores = threads used there, all_cores = tok/s with every core, sweep, kind, sample, ms_per_step, steps)."""
    from oracle import binding as ob
    nproc = os.cpu_count() or 1
    kind = "reference"
    os.environ.setdefault("OMP_WAIT_POLICY", "active")
    os.environ.setdefault("OMP_NUM_THREADS", str(nproc))
    if ob.RefLib.available("fast"):
        ref = ob.RefLib("fast")
        t = time.time()
        m = ref.open(path, seq_len)
        log(f"[bench] reference model_create {time.time() - t:.1f}s ({os.path.basename(ref.path)})")
        fwd = lambda tok, pos: ref.lib.forward(m, tok, pos)  # noqa: E731
        close = lambda: ref.close(m)  # noqa: E731
        counts = sorted({1 << i for i in range(nproc.bit_length()) if (1 << i) <= nproc} | {nproc})
        if not _omp_set_threads(nproc):
            counts = [nproc]
    else:  # the oracle port always exists (serial C)
        kind = "port"
        orc = ob.Oracle()
        om = orc.open(path, seq_len)
        fwd = lambda tok, pos: orc.lib.orc_forward(om.h, tok, pos)  # noqa: E731
        close = om.close
        counts = [1]
    t_begin = time.time()
    sweep, pos = {}, pos0
    for n in reversed(counts):  # many threads first: if the budget runs out the slow single-thread points are the ones skipped
        if kind == "reference":
            _omp_set_threads(n)
        fwd(7, pos); pos += 1  # warm (page cache, thread team)
        best = None
        for _ in range(2):
            t0 = time.time()
            fwd(7, pos); pos += 1
            dt = time.time() - t0
            best = dt if best is None else min(best, dt)
            if time.time() - t_begin > 0.45 * budget_s:
                break
        sweep[n] = 1.0 / best
        if time.time() - t_begin > 0.45 * budget_s:
            break
    n_best = max(sweep, key=sweep.get)
    if kind == "reference":
        _omp_set_threads(n_best)
    per = 1.0 / sweep[n_best]
    left = max(1.0, budget_s - (time.time() - t_begin))
    w = max(0, min(warmup, int(0.15 * left / per)))
    for i in range(w):
        fwd(7, pos); pos += 1
    k = max(1, min(steps, int(0.7 * left / per)))
    first = pos
    t1 = time.time()
    for i in range(k):
        fwd(7, pos); pos += 1
    dt = time.time() - t1
    close()
    return {"value": k / dt, "unit": "tok/s", "cores": n_best, "kind": kind, "ms_per_step": 1e3 * dt / k, "steps": k,
            "all_cores": {"threads": nproc, "value": sweep.get(nproc)},
            "sweep": {str(n): round(v, 3) for n, v in sorted(sweep.items())},
            "sample": f"{k} forward() calls at pos {first}.. at the best of an OMP thread sweep {sorted(sweep)} "
                      f"(best {n_best} threads; all {nproc} cores: {sweep.get(nproc, float('nan')):.2f} tok/s), same .bin, "
                      f"{ob.cpu_model()}, OMP_WAIT_POLICY={os.environ.get('OMP_WAIT_POLICY')}"}


def measured_int8_peak(ql):
    """Dense int8 tensor throughput measured on this device (tcgen05.mma.kind::i8 M128 N256 K32 back to back on resident
    operands, csrc/prefill_gemm.cu:k_int8_peak), tera-ops/s; the nominal 4.5 POPS only if the measurement fails."""
    try:
        ql.lib.qwen_cuda_int8_peak.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_float)]
        t = C.c_float(0)
        if ql.lib.qwen_cuda_int8_peak(20000, 3, C.byref(t)) == 0 and t.value > 0:
            return float(t.value), "measured in this run (qwen_cuda_int8_peak: tcgen05.mma.kind::i8 M128 N256 K32, no loads, no epilogue)"
    except Exception:
        pass
    return 4500.0, "nominal dense int8 (measurement failed)"


def single_gpu_decode(pkg, ql, shape_name, ctx, K, W):
    """Device-timed decode of `shape_name` on ONE GPU (this process's current device): the same-workload base of the
    tensor-parallel lines (the scaling curve must divide like by like)."""
    path = ensure_ckpt(pkg, shape_name)
    t = time.time()
    gm = ql.open(path, ctx + W + K + 8)
    create_s = time.time() - t
    ms, _ = gm.time_decode(7, ctx, K, W)
    gm.close()
    shape = pkg.checkpoint.SHAPES[shape_name]
    bytes_tok = float(np.mean([shape.decode_bytes(ctx + W + i) for i in range(K)]))
    peak, _ = measured_peak()
    return {"workload": f"{shape_name}-decode-ctx{ctx}", "n_gpus": 1, "value": K / (ms / 1e3), "unit": "tok/s", "ms_per_step": ms / K,
            "hbm_frac": bytes_tok * K / (ms / 1e3) / 1e9 / peak, "model_create_s": create_s}


def _timed(fn):
    t0 = time.perf_counter()
    if not fn():
        raise RuntimeError("prefill failed")
    return time.perf_counter() - t0


def cli_generation(pkg, path, shape):
    """SURVEY.md 8f-2 at program level: the reference's own CLI (examples/qwen.c + src/completion.c with
    integration/completion_fast.patch: prompt through forward_prefill, sampling on the device), linked against libqwen3.so
    (oracle/_ref/qwen_b200_fast, built by oracle/Makefile where the reference tree is present). Two runs with the same
    64-token prompt and seed at context 192 and 704: start-up, model load and the prefill cancel in the difference, which is
    512 generated tokens (the synthetic tokenizer has no stop token, so generation runs to the end of the context)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "qwen_b200_fast")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/qwen_b200_fast not built (needs the reference tree at build time)"}
    if not os.path.exists(path + ".tokenizer"):
        pkg.checkpoint.write_tokenizer(path + ".tokenizer", shape.vocab_size)
    prompt = "the quick brown fox jumps over the lazy dog. " * 2
    prompt = prompt[:64]
    times = {}
    for ctx in (192, 704):
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            r = subprocess.run([exe, path, "-m", "completion", "-i", prompt, "-c", str(ctx), "-t", "0.7", "-p", "0.8", "-s", "1"],
                               stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return {"error": "exit %d: %s" % (r.returncode, r.stderr[-300:].decode(errors="replace"))}
            best = dt if best is None else min(best, dt)
        times[ctx] = best
    gen = 704 - 192
    return {"value": gen / (times[704] - times[192]), "unit": "tok/s", "generated_tokens": gen, "prompt_tokens": len(prompt),
            "wall_s": {str(k): v for k, v in times.items()},
            "what": "reference CLI + completion_fast.patch on libqwen3.so, -m completion -t 0.7 -p 0.8: wall-time difference of two "
                    "runs (context 704 vs 192), i.e. model load, CUDA start-up and the prompt prefill cancel"}


def config2_job(pkg, ql):
    """BASELINE config 2 as one job on one GPU: Qwen3-1.7B shape, a 512-token prompt through forward_prefill, then 256 decode
    steps (sampled on the device, token fed back) -- wall time of the whole job from host tokens in to the last token out."""
    shape_name = "1.7b"
    shape = pkg.checkpoint.SHAPES[shape_name]
    path = ensure_ckpt(pkg, shape_name)
    Tp, Tg = 512, 256
    toks = [int(t) for t in np.random.default_rng(1).integers(0, shape.vocab_size, size=Tp)]
    state = [np.uint64(0x9E3779B97F4A7C15)]

    def coin():  # xorshift* as in the reference's sampler
        x = int(state[0])
        x ^= x >> 12
        x ^= (x << 25) & 0xFFFFFFFFFFFFFFFF
        x ^= x >> 27
        state[0] = np.uint64(x)
        return float(np.float32(((x * 0x2545F4914F6CDD1D & 0xFFFFFFFFFFFFFFFF) >> 32 >> 8) / 16777216.0))

    gm = ql.open(path, Tp + Tg + 8)
    try:
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            lg = gm.forward_prefill(toks, 0)
            t1 = time.perf_counter()
            tok = int(np.argmax(lg))
            for i in range(Tg):
                gm.forward_async(tok, Tp + i)
                tok = gm.sample(0.7, 0.8, coin()) or 7
            t2 = time.perf_counter()
            if best is None or t2 - t0 < best[0]:
                best = (t2 - t0, t1 - t0, t2 - t1)
    finally:
        gm.close()
    return {"workload": "1.7b-prefill512+decode256", "job_s": best[0], "prefill_ms": 1e3 * best[1], "prefill_tok_s": Tp / best[1],
            "decode_tok_s": Tg / best[2], "tok_s_overall": (Tp + Tg) / best[0],
            "what": "forward_prefill (logits of the last prompt token to the host) then forward_async + qwen_cuda_sample per token; best of 3"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=256)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tp-base", action="store_true", help="skip the 8B one-GPU figure on the N=1 line")
    ap.add_argument("--no-tp1", action="store_true", help="N > 1: skip the same-workload one-GPU leg")
    ap.add_argument("--path", default="mega", choices=["mega", "ops"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(args.gpus, world)
    workload = args.workload or ("4b-decode-ctx4096" if n_gpus == 1 else "8b-decode-ctx4096")
    shape_name, ctx = WORKLOADS[workload]
    K, W = args.steps, max(args.warmup, 3)
    pkg = entry._pkg()
    shape = pkg.checkpoint.SHAPES[shape_name]
    seq_len = ctx + W + K + 48  # room for the CPU arm's thread sweep (reference forward() trusts pos)
    pos0 = ctx
    base = {"metric": "decode_tokens_per_s", "unit": "tok/s", "n_gpus": n_gpus, "steps": K, "warmup": W,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "data": "synthetic",
            "config": {"workload": workload, "shape": shape_name, "context": ctx, "quant": "Q8_0 (int8 + fp32 scale per 64)",
                       "kv_cache": "fp32", "weights": "random-init .bin in qwen3.c format (mode=fast, seed 1234)",
                       "l2": "weights per token (%.2f GB) exceed the 126 MB L2, no flush needed" % (shape.weight_elements() * 17 / 16 / 1e9)}}

    if args.impl == "reference":
        if rank != 0:
            return 0
        path = ensure_ckpt(pkg, shape_name)
        r = cpu_reference_run(path, seq_len, pos0, K, W)
        line = dict(base, impl="reference", dtype="int8xint8->int32, fp32", value=r["value"], ms_per_step=r["ms_per_step"],
                    steps=r["steps"], gpu_launches=0,
                    cpu_baseline={"value": r["value"], "unit": "tok/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                  "all_cores": r["all_cores"], "sweep": r["sweep"]},
                    e2e={"value": r["value"], "unit": "tok/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
        print(json.dumps(line), flush=True)
        return 0

    # ------------------------------------------------------------------ our arm
    if n_gpus > 1:
        import bench_tp  # tensor-parallel arm lives in its own module
        return bench_tp.run(args, base, pkg, shape, shape_name, ctx, K, W, rank, world)

    ql = pkg.QwenLib()
    if ql.lib.qwen_cuda_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    path = ensure_ckpt(pkg, shape_name)
    t = time.time()
    gm = ql.open(path, seq_len)
    model_create_s = time.time() - t
    log(f"[bench] model_create {model_create_s:.1f}s")
    gm.set_path(0 if args.path == "mega" else 1)

    sampler = ClockSampler(0)
    sampler.start()
    # (1) device-resident: K steps back to back, CUDA events on the launching stream
    ms, launches = gm.time_decode(7, pos0, K, W)
    # (2) end to end through forward(): token/pos in, logits copied to pinned host memory every step
    for i in range(W):
        gm.forward_nocopy(7, pos0 + i)
    t0 = time.perf_counter()
    for i in range(K):
        ptr = gm.forward_nocopy(7, pos0 + W + i)
        if not ptr:
            raise SystemExit("forward failed: " + ql.err())
    e2e_s = time.perf_counter() - t0
    # (2b) generation as the CLI does it, with the sampler on the device: step, then qwen_cuda_sample
    # (temperature 0.7, top-p 0.8, host xorshift coin) returns the token id -- 8 bytes cross PCIe per step, not 608 KB,
    # and no 150k-entry qsort on the host. The sampled token feeds the next step.
    sampled = None
    try:
        rng_state = np.uint64(0x9E3779B97F4A7C15)

        def coin():
            nonlocal rng_state  # reference src/xorshift.c:7-16
            x = int(rng_state)
            x ^= x >> 12
            x ^= (x << 25) & 0xFFFFFFFFFFFFFFFF
            x ^= x >> 27
            rng_state = np.uint64(x)
            return float(np.float32(((x * 0x2545F4914F6CDD1D & 0xFFFFFFFFFFFFFFFF) >> 32 >> 8) / 16777216.0))

        tok, declined = 7, 0
        for i in range(W):
            gm.forward_async(tok, pos0 + i)
            tok = gm.sample(0.7, 0.8, coin()) or 7
        t0 = time.perf_counter()
        for i in range(K):
            gm.forward_async(tok, pos0 + W + i)
            nxt = gm.sample(0.7, 0.8, coin())
            if nxt is None:  # fast path declined: the CLI would fall back to the host sampler for this step
                declined += 1
                nxt = 7
            tok = nxt
        dt = time.perf_counter() - t0
        sampled = {"value": K / dt, "unit": "tok/s", "d2h_bytes_per_step": 8, "h2d_bytes_per_step": 8, "declined_steps": declined,
                   "what": "forward_async + qwen_cuda_sample(temperature 0.7, top_p 0.8), sampled token fed back"}
    except Exception as e:  # never lose the decode number
        sampled = {"error": repr(e)}
    clocks = sampler.stop()
    ql._ok(ql.lib.qwen_cuda_sync(gm.ctx), "sync")

    tok_s = K / (ms / 1e3)
    bytes_tok = float(np.mean([shape.decode_bytes(pos0 + W + i) for i in range(K)]))
    peak, peak_src = measured_peak()
    achieved = bytes_tok * tok_s / 1e9
    line = dict(base, value=tok_s, ms_per_step=ms / K, dtype="int8xint8->int32, fp32", clocks=clocks, gpu_launches=launches,
                e2e={"value": K / e2e_s, "unit": "tok/s", "h2d_bytes_per_step": 8, "d2h_bytes_per_step": shape.vocab_size * 4},
                roofline={"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                          "traffic": None, "peak_source": peak_src, "bytes_per_token": bytes_tok,
                          "kernel": KERNEL_TAG + " (persistent, 1 launch/token)" if args.path == "mega" else "per-op kernels",
                          "frac_of_8TBs_nominal": achieved / 8000.0})
    line["config"]["path"] = args.path
    line["model_create_s"] = model_create_s
    line["sampled_generation"] = sampled
    try:  # dram bytes per launch from this round's ncu capture -- only if it was taken on the kernel variant that just ran
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
        ent = tr.get(workload)
        overridden = [k for k in os.environ if k.startswith("QWEN_MEGA_")]
        if args.path == "mega" and ent and ent.get("kernel") == KERNEL_TAG and not overridden:
            line["roofline"]["traffic"] = ent["traffic"]
            line["roofline"]["traffic_source"] = ("ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch of %s, captured at commit %s "
                                                  "(profiles/r2_traffic.json); not re-measured in this run" % (ent["kernel"], ent.get("commit", "?")))
    except Exception:
        pass
    # (3) prompt prefill on the same model (north_star: prefill tok/s + int8 tensor-pipe fraction): 512 tokens at
    # positions 0..511 through forward_prefill's device entry, host tokens in, no logits copy; best of 3
    try:
        Tp = 512
        toks = [int(t) for t in np.random.default_rng(0).integers(0, shape.vocab_size, size=Tp)]
        gm.prefill_nocopy(toks, 0)
        best = min(_timed(lambda: gm.prefill_nocopy(toks, 0)) for _ in range(3))
        macs = Tp * (shape.weight_elements() - shape.vocab_size * shape.dim)  # layer GEMMs; the classifier runs for the last token only
        peak_i8, peak_i8_src = measured_int8_peak(ql)
        line["prefill"] = {"workload": f"{shape_name}-prefill{Tp}", "tok_s": Tp / best, "ms": 1e3 * best,
                           "int8_tops": 2 * macs / best / 1e12, "peak_tops": peak_i8,
                           "frac": 2 * macs / best / 1e12 / peak_i8,
                           "peak_source": peak_i8_src,
                           "note": ("whole 512-token prefill (GEMMs + attention + norm/quantise passes) over the layer GEMMs' int8 ops; the "
                                    "group-scaled GEMM promotes every 64-wide group to fp32 on the CUDA cores (5 issue cycles per output and "
                                    "group: ceiling ~20 % of the tensor peak for the GEMMs alone), profiles/r2_prefill_summary.md")}
    except Exception as e:  # never lose the decode number
        line["prefill"] = {"error": repr(e)}
    gm.close()
    if workload == "4b-decode-ctx4096" and not args.no_tp_base:
        try:  # (4) the reference's own program on this library (patched generation loops)
            line["cli"] = cli_generation(pkg, path, shape)
        except Exception as e:
            line["cli"] = {"error": repr(e)}
        try:  # (5) BASELINE config 2 as one job
            line["config2"] = config2_job(pkg, ql)
        except Exception as e:
            line["config2"] = {"error": repr(e)}
    if workload == "4b-decode-ctx4096" and not args.no_tp_base:
        # the tensor-parallel lines (N >= 2) run the 8B shape: its one-GPU figure, so that the 1 -> 8 curve divides like by like
        try:
            line["tp_base"] = single_gpu_decode(pkg, ql, "8b", 4096, K, W)
        except Exception as e:
            line["tp_base"] = {"error": repr(e)}
    if not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(path, seq_len, pos0, 6, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": "tok/s", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"],
                                    "all_cores": r["all_cores"], "sweep": r["sweep"]}
        except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
            line["cpu_baseline"] = {"value": None, "unit": "tok/s", "cores": 0, "kind": "unavailable", "sample": repr(e)}
    print(json.dumps(line), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
