"""Throughput of the tcgen05 prefill GEMM on projection-sized problems (device time, best of reps)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as e
pkg = e._pkg(); ql = pkg.QwenLib()
rng = np.random.default_rng(0)
for name, d, n, T in (("4B qkv", 6144, 2560, 512), ("4B w1/w3", 19456, 2560, 512), ("4B w2", 2560, 9728, 512), ("1.7B qkv T=2048", 4096, 2048, 2048)):
    xq = rng.integers(-127, 128, size=(T, n), dtype=np.int8)
    xs = rng.uniform(1e-3, 1e-1, size=(T, n // 64)).astype(np.float32)
    wq = rng.integers(-127, 128, size=d * n, dtype=np.int8)
    ws = rng.uniform(1e-4, 1e-2, size=d * n // 64).astype(np.float32)
    out, _, ms = ql.matmul_batch(xq, xs, wq, ws, n, d, T, reps=5)
    ops = 2.0 * T * d * n
    print(f"{name:18s} d={d:6d} n={n:5d} T={T:5d}: {ms*1e3:8.1f} us  {ops/ms/1e9:8.1f} TOPS (int8 dense)  {T/(ms/1e3)/1e6:6.2f} M tok-rows/s", flush=True)
