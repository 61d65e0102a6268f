"""Long-prompt cross-check of the prefill kernel variants: the same 2048-token prompt through forward_prefill with the round-2
kernels and with the round-1 kernels (QWEN_ATTN_V=1 QWEN_GEMM_V=1, separate process); prints the largest logit difference."""
import os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
shape = sys.argv[1] if len(sys.argv) > 1 else "1.7b"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
if len(sys.argv) > 3:  # child: write the logits
    import __graft_entry__ as e
    pkg = e._pkg(); ql = pkg.QwenLib()
    path = pkg.checkpoint.ensure_checkpoint("/tmp/qwen3_b200_ckpt", shape, seed=1234, mode="fast")
    toks = [int(t) for t in np.random.default_rng(0).integers(0, 1000, size=T)]
    with ql.open(path, T + 64) as gm:
        lg = gm.forward_prefill(toks, 0)
        lg2 = gm.forward(int(np.argmax(lg)), T)  # one decode step on the prefilled cache
    np.save(sys.argv[3], np.stack([lg, lg2]))
    sys.exit(0)
outs = []
variants = [("round-2 kernels", {}), ("round-1 kernels", {"QWEN_ATTN_V": "1", "QWEN_GEMM_V": "1"}), ("round-1 GEMM only", {"QWEN_GEMM_V": "1"})]
for name, env in variants:
    f = f"/tmp/prefill_cmp_{len(outs)}.npy"
    subprocess.run([sys.executable, __file__, shape, str(T), f], env=dict(os.environ, **env), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    outs.append(np.load(f))
a = outs[0]
for (name, _), b in zip(variants[1:], outs[1:]):
    for i, what in enumerate(("last prompt token", "first decode step")):
        d = np.abs(a[i] - b[i])
        print(f"{shape} T={T} round-2 vs {name}, {what}: max |dlogit| {d.max():.3e} (logit std {b[i].std():.3f}), argmax {int(a[i].argmax())} / {int(b[i].argmax())}", flush=True)
