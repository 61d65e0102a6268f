"""GPU parity tests (pytest -m gpu): the CUDA path, called through the C ABI with the
reference's own function names, against the pinned CPU oracle on identical inputs.

Bars (north_star):
  * q8_quantize / q8_dequantize / every Q8_0 group's int32 dot / rotary: BIT-EXACT;
  * fp32 ops and logits: |a-b| <= ATOL + RTOL*|b| with RTOL 1e-3, ATOL 1e-2;
  * greedy decoding: identical token sequence.
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-2
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def qlib(pkg):
    pkg.build.build()
    q = pkg.QwenLib()
    assert q.lib.qwen_cuda_device_count() > 0, "GPU tests need a CUDA device"
    return q


@pytest.fixture(scope="module")
def gold():
    return dict(np.load(os.path.join(GOLD, "micro_golden.npz")))


def close(a, b, rtol=RTOL, atol=ATOL):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    bad = np.abs(a - b) > atol + rtol * np.abs(b)
    assert not bad.any(), f"{bad.sum()} of {bad.size} outside tolerance; max abs diff {np.abs(a - b).max():.3e}"


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape
    assert np.array_equal(a.view(np.uint8), b.view(np.uint8)), f"not bit-identical, max abs diff {np.abs(a.astype(np.float64) - b).max()}"


# ---------------------------------------------------------------- ops
@pytest.mark.parametrize("n", [64, 1024, 2560, 9728, 25600])
def test_quantize_bit_exact(qlib, oracle, n):
    rng = np.random.default_rng(n)
    for scale in (1e-3, 1.0, 300.0):
        x = (rng.standard_normal(n) * scale).astype(np.float32)
        x[: min(n, 64)] = 0  # one all-zero group
        q, s = qlib.q8_quantize(x)
        oq, os_ = oracle.q8_quantize(x)
        same(q, oq)
        same(s, os_)


def test_quantize_rounding_ties_and_golden_edges(qlib, oracle, gold):
    q, s = qlib.q8_quantize(gold["quant_x"])
    same(q, gold["quant_q"])
    same(s, gold["quant_s"])
    same(qlib.q8_dequantize(q, s), gold["dequant_x"])
    # exact .5 ties after division: x = (k + 0.5) * scale with scale = 127/127 = 1
    x = np.zeros(64, np.float32)
    x[0] = 127.0
    x[1:40] = np.arange(39, dtype=np.float32) + 0.5
    x[40:63] = -(np.arange(23, dtype=np.float32) + 0.5)
    q, s = qlib.q8_quantize(x)
    oq, os_ = oracle.q8_quantize(x)
    same(q, oq)
    assert q[1] == 1 and q[40] == -1  # half away from zero, not half-to-even


def test_quantize_tail_ignored_and_empty(qlib, oracle):
    x = np.arange(100, dtype=np.float32)
    q, s = qlib.q8_quantize(x)
    oq, os_ = oracle.q8_quantize(x)
    same(q[:64], oq[:64])
    assert not q[64:].any()  # tail n % 64 untouched, like the reference
    q0, s0 = qlib.q8_quantize(np.zeros(0, np.float32))
    assert q0.size == 0


@pytest.mark.parametrize("n,d", [(64, 1), (320, 37), (1024, 130), (2560, 64), (9728, 33), (4096, 257)])
def test_matmul_group_dots_bit_exact_and_output_close(qlib, oracle, n, d):
    rng = np.random.default_rng(n * 1000 + d)
    xq = rng.integers(-127, 128, size=n, dtype=np.int8)
    wq = rng.integers(-127, 128, size=n * d, dtype=np.int8)
    wq[:64] = 127
    xq[:64] = -127  # extreme group: dot = -64*127*127
    xs = rng.uniform(1e-3, 1e-1, size=n // 64).astype(np.float32)
    ws = rng.uniform(1e-4, 1e-2, size=n * d // 64).astype(np.float32)
    dots = qlib.group_dots(xq, wq, n, d)
    same(dots, oracle.group_dots(xq, wq, n, d))
    assert dots[0, 0] == -64 * 127 * 127
    out = qlib.matmul(xq, xs, wq, ws, n, d)
    ref = oracle.matmul(xq, xs, wq, ws, n, d)
    close(out, ref, rtol=1e-5, atol=1e-5 * float(np.abs(ref).max()))


def test_matmul_golden(qlib, gold):
    n, d = map(int, gold["mm_nd"])
    out = qlib.matmul(gold["mm_xq"], gold["mm_xs"], gold["mm_wq"], gold["mm_ws"], n, d)
    close(out, gold["mm_out"], rtol=1e-5, atol=1e-5 * float(np.abs(gold["mm_out"]).max()))


def test_elementwise_ops(qlib, oracle, gold):
    for tag in ("big", "head"):
        close(qlib.rmsnorm(gold[f"rms_{tag}_x"], gold[f"rms_{tag}_w"]), gold[f"rms_{tag}_out"], rtol=1e-5, atol=1e-6)
    close(qlib.softmax(gold["softmax_x"]), gold["softmax_out"], rtol=1e-5, atol=1e-9)
    for pos in (0, 1, 777, 4095, 32767):
        same(qlib.rotary(gold["rot_x"], 128, pos), gold[f"rot_out_{pos}"])  # host libm angles -> exact
    close(qlib.swiglu(gold["swiglu_x1"], gold["swiglu_x3"]), gold["swiglu_out"], rtol=1e-5, atol=1e-7)
    silu = np.array([qlib.lib.silu(float(v)) for v in gold["silu_x"]], np.float32)
    sig = np.array([qlib.lib.sigmoid(float(v)) for v in gold["silu_x"]], np.float32)
    close(silu, gold["silu_out"], rtol=1e-6, atol=1e-9)
    close(sig, gold["sigmoid_out"], rtol=1e-6, atol=1e-12)
    big = np.random.default_rng(5).standard_normal(151936).astype(np.float32) * 8
    close(qlib.softmax(big), oracle.softmax(big), rtol=1e-4, atol=1e-10)  # sampler-sized (sampler.c:196)
    x = np.random.default_rng(6).standard_normal(2560).astype(np.float32)
    inplace = x.copy()
    qlib.lib.rmsnorm(inplace.ctypes.data_as(pkg_fp()), inplace.ctypes.data_as(pkg_fp()),
                     np.ones(2560, np.float32).ctypes.data_as(pkg_fp()), 2560)
    close(inplace, oracle.rmsnorm(x, np.ones(2560, np.float32)), rtol=1e-5, atol=1e-6)


def pkg_fp():
    import ctypes
    return ctypes.POINTER(ctypes.c_float)


# ---------------------------------------------------------------- whole forward
#
# What "logits within rtol 1e-3 / atol 1e-2" can and cannot mean for this model family.
# Activations are re-quantised to int8 before each of the 7L+1 GEMVs (q8.c:5-30). Any fp32
# reordering upstream (a tree sum instead of the reference's serial loop) moves a value by
# ~1 ulp, and about 6e-6 of all elements then sit on the other side of a rounding boundary:
# one int8 code flips. A flip in the classifier input alone moves logits by
# |cls[v][j]| * scale ~ 3e-2. The reference is subject to exactly this against ITSELF: its
# README build (-Ofast -fopenmp, 4 threads) differs from its serial -O2 build by up to 0.19
# logit units on the 0.6B shape (measured, DESIGN.md "Parity"). So the tests below require
#   (1) every step: same argmax as the oracle unless the oracle's own top-2 margin < 2*ATOL;
#   (2) every step: max |dlogit| <= NOISE_CAP * std(logits)  (a real bug is O(1), not O(1e-2));
#   (3) every step: max |dlogit| <= 3 x max(flip unit, the reference's own worst step), where
#       the flip unit is what ONE flipped int8 code in the classifier input does to a logit
#       (largest classifier weight x largest activation scale, both read from the oracle) and
#       "the reference's own worst" is its -Ofast/OpenMP build against its strict build on
#       the same tokens, run beside the GPU in the same test when oracle/_ref is present;
#   (4) on the reference-exported golden checkpoint (12 steps, no flip) the tolerance holds
#       outright (test_forward_golden_micro).
# Which steps flip is luck (the reference's two builds flip on different steps in different
# runs), so the share of steps meeting the tolerance is printed, not compared.
NOISE_CAP = 0.05


def within(a, b, rtol=RTOL, atol=ATOL):
    return bool((np.abs(a.astype(np.float64) - b) <= atol + rtol * np.abs(b)).all())


def parity_run(qlib, oracle, path, tokens, seq_len, path_sel, with_ref=True):
    """Teacher-forced run of `tokens` on GPU, strict oracle and (if present) the reference's own
    fast build. Returns per-step stats and the GPU / oracle KV caches."""
    from oracle.binding import RefLib
    ref = fm = None
    if with_ref and RefLib.available("fast"):
        os.environ.setdefault("OMP_NUM_THREADS", "4")
        ref = RefLib("fast")
        fm = ref.open(path, seq_len)
    stats = []
    with qlib.open(path, seq_len) as gm, oracle.open(path, seq_len) as om:
        gm.set_path(path_sel)
        for pos, t in enumerate(tokens):
            lg, lo = gm.forward(int(t), pos), om.forward(int(t), pos)
            lf = ref.forward(fm, int(t), pos) if ref else None
            top, margin = oracle.argmax(lo)
            groups = om.p.dim // 64
            sx = float(np.ctypeslib.as_array(om.p.as_, shape=(groups,)).max())  # classifier-input scales
            if pos == 0:
                ncls = om.p.vocab_size * om.p.dim // 64
                w_max = 127.0 * float(np.ctypeslib.as_array(om.p.cls.s, shape=(ncls,)).max())
            stats.append(dict(pos=pos, flip=w_max * sx, gpu_ok=within(lg, lo), gpu_max=float(np.abs(lg - lo).max()),
                              ref_ok=within(lf, lo) if ref else None,
                              ref_max=float(np.abs(lf - lo).max()) if ref else None,
                              std=float(lo.std()), argmax_same=int(np.argmax(lg)) == top, margin=margin))
        gk = [gm.kv_read(l, 0, len(tokens)) for l in range(gm.p.n_layers)]
        ok, ov = om.kv()
    if ref:
        ref.close(fm)
    return stats, gk, (ok, ov)


def check_parity_stats(stats):
    n = len(stats)
    for s in stats:
        assert s["argmax_same"] or s["margin"] < 2 * ATOL, f"greedy token differs at pos {s['pos']} (margin {s['margin']})"
        assert s["gpu_max"] <= NOISE_CAP * s["std"], f"pos {s['pos']}: |dlogit| {s['gpu_max']:.3e} vs std {s['std']:.3f}"
    gpu_ok = sum(s["gpu_ok"] for s in stats)
    gpu_worst = max(s["gpu_max"] for s in stats)
    flip = max(s["flip"] for s in stats)
    msg = f"steps within rtol 1e-3/atol 1e-2: gpu {gpu_ok}/{n}, worst {gpu_worst:.3e}, one-flip unit {flip:.3e}"
    ref_worst = 0.0
    if stats[0]["ref_ok"] is not None:
        ref_ok = sum(s["ref_ok"] for s in stats)
        ref_worst = max(s["ref_max"] for s in stats)
        msg += f"; reference -Ofast/OpenMP vs its strict build: {ref_ok}/{n}, worst {ref_worst:.3e}"
    print("\n[parity] " + msg)
    assert gpu_worst <= 3 * max(flip, ref_worst), msg
    return msg


@pytest.mark.parametrize("path_sel", [1, 0])
def test_forward_golden_micro(qlib, oracle, gold, path_sel):
    """Short run on the reference-exported golden checkpoint: tolerance holds outright."""
    toks = gold["tokens"]
    with qlib.open(os.path.join(GOLD, "micro.bin")) as gm:
        gm.set_path(path_sel)
        for pos, t in enumerate(toks):
            close(gm.forward(int(t), pos), gold["logits"][pos])
        for l in range(gm.p.n_layers):
            k, v = gm.kv_read(l, 0, len(toks))
            close(k, gold["k_cache"][l], rtol=1e-4, atol=1e-4)
            close(v, gold["v_cache"][l], rtol=1e-4, atol=1e-4)
        att = gm.attention(1, len(toks) - 1, gold["att_q"])
        close(att, gold["att_out"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("shape", ["tiny", "tiny-untied", "small"])
@pytest.mark.parametrize("path_sel", [1, 0])
def test_forward_logits_and_kv_match_oracle(qlib, oracle, pkg, ckpt_dir, shape, path_sel):
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, shape, seed=11)
    V = pkg.checkpoint.SHAPES[shape].vocab_size
    toks = np.random.default_rng(2).integers(0, V, size=40)
    stats, gk, (ok, ov) = parity_run(qlib, oracle, path, toks, 64, path_sel)
    check_parity_stats(stats)
    for l, (k, v) in enumerate(gk):  # a flipped code moves a K/V row by ~1e-2; a bug by O(1)
        close(k, ok[l, : len(toks)], rtol=2e-2, atol=2e-2)
        close(v, ov[l, : len(toks)], rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize("shape", ["4b-l2", "8b-l2", "32b-l2"])
def test_real_layer_shapes_decode_and_prefill(qlib, oracle, pkg, ckpt_dir, shape):
    """Two layers of the real Qwen3-4B / R1-Qwen3-8B / Qwen3-32B layer shapes (GQA ratio 4, 4, 8; hidden 9728,
    12288, 25600 -- one w2 row per ring tile at 32B): the persistent kernel at a short and at a long context
    (injected KV rows, so the split-KV attention spans many CTAs and tiles) and forward_prefill, against the oracle."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, shape, seed=11)
    sh = pkg.checkpoint.SHAPES[shape]
    rng = np.random.default_rng(9)
    ctx = 700
    with qlib.open(path, ctx + 16) as gm, oracle.open(path, ctx + 16) as om:
        # identical random K/V rows on both sides for positions 0..ctx-1
        k = (rng.standard_normal((sh.n_layers, ctx + 16, sh.kv_dim)) * 0.5).astype(np.float32)
        v = (rng.standard_normal((sh.n_layers, ctx + 16, sh.kv_dim)) * 0.5).astype(np.float32)
        om.set_kv(k, v)
        for l in range(sh.n_layers):
            gm.kv_write(l, 0, k[l, :ctx], v[l, :ctx])
        tok, within = 5, 0
        for pos in range(ctx, ctx + 4):
            lg, lo = gm.forward(tok, pos), om.forward(tok, pos)
            assert np.abs(lg - lo).max() < 0.05 * max(1.0, lo.std()), (pos, np.abs(lg - lo).max())
            nxt, margin = oracle.argmax(lo)
            assert int(np.argmax(lg)) == nxt or margin < 2e-2
            within += int(not (np.abs(lg - lo) > 1e-2 + 1e-3 * np.abs(lo)).any())
            tok = nxt
        assert within >= 1, "no step met rtol 1e-3 / atol 1e-2 outright (one flipped int8 code explains some, not all)"
    prompt = [int(t) for t in rng.integers(0, sh.vocab_size, size=40)]
    with qlib.open(path, 64) as gm, oracle.open(path, 64) as om:
        for pos, t in enumerate(prompt):
            lo = om.forward(t, pos)
        lg = gm.forward_prefill(prompt, 0)
        assert np.abs(lg - lo).max() < 0.05 * max(1.0, lo.std()), np.abs(lg - lo).max()
        nxt, margin = oracle.argmax(lo)
        assert int(np.argmax(lg)) == nxt or margin < 2e-2


@pytest.mark.parametrize("path_sel", [1, 0])
def test_greedy_256_tokens_identical(qlib, oracle, pkg, ckpt_dir, path_sel):
    """north_star: greedy decoding gives an identical 256-token sequence. Ties are only excused
    when the oracle's own top-1/top-2 margin is below the logits tolerance."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "small", seed=11)
    n = 256
    with qlib.open(path, n + 1) as gm, oracle.open(path, n + 1) as om:
        gm.set_path(path_sel)
        tok_g = tok_o = 17
        seq_g, seq_o, min_margin = [], [], 1e9
        for pos in range(n):
            lg, lo = gm.forward(tok_g, pos), om.forward(tok_o, pos)
            tok_o, margin = oracle.argmax(lo)
            tok_g = int(np.argmax(lg))
            min_margin = min(min_margin, margin)
            seq_g.append(tok_g)
            seq_o.append(tok_o)
            if tok_g != tok_o:
                assert margin < 2 * ATOL, f"diverged at {pos} with oracle margin {margin}"
                tok_g = tok_o  # a genuine tie: follow the oracle and keep comparing
        assert seq_g == seq_o, f"sequences differ (min oracle margin {min_margin})"
    # the device-resident greedy chain must produce the same tokens as forward()+argmax
    with qlib.open(path, n + 1) as gm:
        gm.set_path(path_sel)
        chain = gm.decode_greedy(17, 0, 64)
        assert list(chain) == seq_o[:64]


def test_attention_long_context_with_injected_kv(qlib, oracle, pkg, ckpt_dir):
    """Config-3 style step without hours of CPU prefill: identical random K/V injected on both
    sides, then attention() and one decode step at pos 1000 are compared."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "small", seed=11)
    sh = pkg.checkpoint.SHAPES["small"]
    S, pos = 1024, 1000
    rng = np.random.default_rng(8)
    k = rng.standard_normal((sh.n_layers, S, sh.kv_dim)).astype(np.float32)
    v = rng.standard_normal((sh.n_layers, S, sh.kv_dim)).astype(np.float32)
    for path_sel in (1, 0):
        with qlib.open(path, S) as gm, oracle.open(path, S) as om:
            gm.set_path(path_sel)
            om.set_kv(k, v)
            for l in range(sh.n_layers):
                gm.kv_write(l, 0, k[l], v[l])
                rk, rv = gm.kv_read(l, 0, S)  # layout round trip: [pos][kv_dim] <-> [kv_head][pos][128]
                same(rk, k[l])
                same(rv, v[l])
            q = rng.standard_normal(sh.proj_dim).astype(np.float32)
            ref = oracle.attention(q, k[2], v[2], sh.n_heads, sh.n_kv_heads, 128, S, pos)
            close(gm.attention(2, pos, q), ref, rtol=1e-4, atol=1e-5)
            lg, lo = gm.forward(5, pos), om.forward(5, pos)
            assert int(np.argmax(lg)) == int(np.argmax(lo))
            assert np.abs(lg - lo).max() <= NOISE_CAP * lo.std()


def test_layered_intermediates_match_oracle_trace(qlib, oracle, pkg, ckpt_dir):
    """Per-layer intermediates (attention output, SwiGLU output) of both GPU paths against the
    oracle's trace, layer by layer via the run-first-n-layers debug knob."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "tiny-untied", seed=11)
    sh = pkg.checkpoint.SHAPES["tiny-untied"]
    toks = [3, 77, 200]
    with oracle.open(path, 16, trace=True) as om:
        for pos, t in enumerate(toks):
            om.forward(t, pos)
        tr = om.trace()
    for path_sel in (1, 0):
        for n in range(1, sh.n_layers + 1):
            with qlib.open(path, 16) as gm:
                gm.set_path(path_sel)
                for pos, t in enumerate(toks[:-1]):
                    gm.forward(t, pos)
                gm.set_layers(n)
                gm.forward(toks[-1], len(toks) - 1)
                close(gm.debug_read("att", sh.proj_dim), tr["att_out"][n - 1], rtol=1e-3, atol=1e-4)
                close(gm.debug_read("h", sh.hidden_dim), tr["h"][n - 1], rtol=1e-3, atol=1e-4)


def test_forward_rejects_out_of_range_pos(qlib, pkg, ckpt_dir):
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "tiny", seed=11)
    with qlib.open(path, 8) as gm:
        assert gm.p.seq_len == 8
        assert not gm.forward_nocopy(1, 8)
        assert not gm.forward_nocopy(1, -1)
        assert not gm.forward_nocopy(10 ** 6, 0)
        assert gm.forward_nocopy(1, 7)
    with qlib.open(path, 0) as gm:  # 0 = keep the header's value (model.c:74-76)
        assert gm.p.seq_len == pkg.checkpoint.SHAPES["tiny"].seq_len
    with qlib.open(path, 10 ** 6) as gm:  # larger than the header: ignored
        assert gm.p.seq_len == pkg.checkpoint.SHAPES["tiny"].seq_len


def test_06b_shape_logits_match_oracle(qlib, oracle, pkg, ckpt_dir):
    """Config 1's shape (Qwen3-0.6B, 28 layers, vocab 151936) against the oracle and against the
    reference's own self-noise for a short teacher-forced run."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "0.6b", seed=1234, mode="gauss")
    toks = [11, 4711, 151935, 0, 90210, 7, 7, 1234]
    stats, _, _ = parity_run(qlib, oracle, path, toks, 32, 0)
    check_parity_stats(stats)


def test_tensor_parallel_2gpu_matches_single_gpu(qlib, pkg, ckpt_dir):
    """TP=2 (one process per GPU) against the oracle on the same checkpoints and tokens, on the persistent kernel
    with the all-reduce fused as NVLink peer stores AND on the per-op + NCCL path: same greedy tokens on every rank
    and path, logits within the flip-noise bound, bit-identical across ranks. Needs 2 visible GPUs (gpurun --gpus 2)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    # small: 2 kv heads per rank; 8b-l2: the real 8B layer shape; tiny: ONE kv head per rank (what TP=8 looks like on 8 kv heads)
    paths = [pkg.checkpoint.ensure_checkpoint(ckpt_dir, name, seed=11) for name in ("small", "8b-l2", "tiny")]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531",
                          os.path.join(root, "tests", "tp_gpu_worker.py")] + paths,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert out.returncode == 0 and "TP_GPU_OK" in out.stdout, out.stdout[-3000:]


def test_quantize_fast_path_bit_exact_on_many_values(qlib, oracle):
    """4M values, half of the groups seeded with values one ulp either side of k+0.5 after division:
    the device quantiser (IEEE division + roundf) must give the oracle's codes bit for bit. (A
    reciprocal-multiply variant with an exact fallback was also bit-exact here but slower.)"""
    rng = np.random.default_rng(123)
    n = 64 * 65536
    x = rng.standard_normal(n).astype(np.float32) * rng.choice([1e-4, 0.3, 7.0, 900.0], size=n).astype(np.float32)
    g = x.reshape(-1, 64)
    # half the groups: put values exactly on / next to rounding boundaries of that group's scale
    amax = np.abs(g).max(axis=1)
    scale = (amax / np.float32(127.0)).astype(np.float32)
    ks = rng.integers(-126, 126, size=g.shape).astype(np.float32) + np.float32(0.5)
    near = (ks * scale[:, None]).astype(np.float32)
    near = np.nextafter(near, np.where(rng.random(g.shape) < 0.5, np.float32(np.inf), np.float32(-np.inf))).astype(np.float32)
    mask = (np.arange(g.shape[0]) % 2 == 0)[:, None] & (np.abs(near) < amax[:, None])
    mask[:, 0] = False
    g = np.where(mask, near, g)
    x = np.ascontiguousarray(g.reshape(-1), np.float32)
    q, s = qlib.q8_quantize(x)
    oq, os_ = oracle.q8_quantize(x)
    same(s, os_)
    same(q, oq)
    # the quantiser fused into the persistent decode kernel (x * rcp(scale) candidate, exact division within 1e-3
    # of a rounding boundary) on the same adversarial values, plus denormal / huge / zero groups
    q, s = qlib.quantize_fused(x)
    same(s, os_)
    same(q, oq)
    edge = np.zeros(64 * 6, np.float32)
    edge[64:128] = np.float32(1e-41) * rng.standard_normal(64).astype(np.float32)        # denormal group: scale underflows
    edge[128:192] = np.float32(3e38) * rng.uniform(-1, 1, 64).astype(np.float32)          # near FLT_MAX
    edge[192:256] = np.float32(1.17549435e-38) * rng.uniform(-4, 4, 64).astype(np.float32)
    edge[256:320] = np.arange(64, dtype=np.float32) - 31.5                               # exact halves after scaling
    edge[320:384] = rng.standard_normal(64).astype(np.float32)
    q, s = qlib.quantize_fused(edge)
    oq, os_ = oracle.q8_quantize(edge)
    same(s, os_)
    same(q, oq)


@pytest.mark.parametrize("n,d,T", [(64, 8, 1), (320, 300, 200), (2560, 384, 256), (1024, 128, 130)])
def test_prefill_matmul_batch_bit_identical_to_reference_matmul(qlib, oracle, n, d, T):
    """The tcgen05 (kind::i8) group-scaled GEMM computes T tokens at once with the reference's own
    arithmetic: exact int32 group dots, ((float) dot * ws) * xs, fp32 fold in group order. Every
    output must equal the oracle's per-token matmul BIT FOR BIT; ragged T and d exercise the TMA
    out-of-bounds fill."""
    rng = np.random.default_rng(n + d + T)
    xq = rng.integers(-127, 128, size=(T, n), dtype=np.int8)
    xs = rng.uniform(1e-3, 1e-1, size=(T, n // 64)).astype(np.float32)
    wq = rng.integers(-127, 128, size=d * n, dtype=np.int8)
    ws = rng.uniform(1e-4, 1e-2, size=d * n // 64).astype(np.float32)
    xq[0, :64] = 127
    wq[:64] = -127  # extreme group
    out, dots, _ = qlib.matmul_batch(xq, xs, wq, ws, n, d, T, want_dots=True)
    for t in (0, T // 2, T - 1):
        same(dots[t], oracle.group_dots(xq[t], wq, n, d))
        same(out[t], oracle.matmul(xq[t], xs[t], wq, ws, n, d))
    assert dots[0, 0, 0] == -64 * 127 * 127
    # all tokens, cheaply: the fold is reproducible from the integer dots alone
    acc = np.zeros((T, d), np.float32)
    w2 = ws.reshape(d, -1)
    for g in range(n // 64):
        acc = acc + (dots[:, :, g].astype(np.float32) * w2[None, :, g]) * xs[:, None, g]
    same(out, acc)


@pytest.mark.parametrize("shape_name,n_prompt", [("tiny", 5), ("tiny-untied", 37), ("small", 150), ("small", 600)])  # 600: two chunks
def test_prefill_matches_reference_token_by_token(qlib, oracle, pkg, ckpt_dir, shape_name, n_prompt):
    """forward_prefill() (tcgen05 GEMMs + batched ops + causal attention, csrc/prefill.cu) must leave the
    device in the state n forward() calls of the reference leave it in: the last token's logits within
    rtol 1e-3 / atol 1e-2 (argmax equal unless the oracle's own margin is below 2*atol), every KV-cache
    row within tolerance, and decoding from the prefilled cache follows the oracle's greedy tokens."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, shape_name, seed=11)
    sh = pkg.checkpoint.SHAPES[shape_name]
    rng = np.random.default_rng(n_prompt)
    prompt = [int(t) for t in rng.integers(0, sh.vocab_size, size=n_prompt)]
    S = n_prompt + 12
    with qlib.open(path, S) as gm, oracle.open(path, S) as om:
        lo = None
        for pos, tok in enumerate(prompt):
            lo = om.forward(tok, pos)
        lg = gm.forward_prefill(prompt, 0)
        bad = np.abs(lg - lo) > 1e-2 + 1e-3 * np.abs(lo)
        if n_prompt <= 200:
            assert bad.mean() < 1e-3 and np.abs(lg - lo).max() < 0.05 * max(1.0, lo.std()), np.abs(lg - lo).max()
        else:  # hundreds of re-quantised K/V rows behind the last token: one-code flips add up; a bug is O(1)
            assert np.abs(lg - lo).max() < 0.1 * max(1.0, lo.std()), np.abs(lg - lo).max()
        nxt, margin = oracle.argmax(lo)
        assert int(np.argmax(lg)) == nxt or margin < 2e-2
        ok, ov = om.kv()  # [L][seq_len][kv_dim]
        for layer in range(sh.n_layers):
            k, v = gm.kv_read(layer, 0, n_prompt)
            for got, want in ((k.reshape(n_prompt, -1), ok[layer, :n_prompt]), (v.reshape(n_prompt, -1), ov[layer, :n_prompt])):
                if layer == 0:  # nothing but the first norm + quantise + exact GEMM (+ norm/RoPE) in front of it
                    close(got, want, rtol=1e-4, atol=1e-4)
                else:  # deeper rows see re-quantised activations: a rare one-code flip moves a few values
                    d = np.abs(got - want)
                    # a flipped code in one row's input moves that whole row (1 row of 37 = 2.7 % of the values); the chunk attention
                    # itself is pinned at op tolerance by test_prefill_attention_kernels_match_reference_attention
                    share = 3e-2 if n_prompt <= 200 else 5e-2  # long prompts: flips in earlier rows feed every later row's attention
                    assert (d > 2e-3 + 2e-3 * np.abs(want)).mean() < share and d.max() < 0.05 * max(1.0, want.std()), (layer, d.max())
        tok = nxt
        for step in range(8):  # decode from the prefilled cache with the persistent kernel
            lg, lo = gm.forward(tok, n_prompt + step), om.forward(tok, n_prompt + step)
            nxt, margin = oracle.argmax(lo)
            assert int(np.argmax(lg)) == nxt or margin < 2e-2, step
            assert np.abs(lg - lo).max() < 0.05 * max(1.0, lo.std())
            tok = nxt


@pytest.mark.parametrize("n_heads,n_kv_heads,pos0,T", [(4, 2, 0, 70), (8, 2, 37, 200), (2, 2, 130, 64), (16, 2, 0, 129), (3, 1, 5, 1), (4, 2, 700, 70), (2, 1, 0, 512)])
@pytest.mark.parametrize("variant", [2, 1])
def test_prefill_attention_kernels_match_reference_attention(qlib, oracle, n_heads, n_kv_heads, pos0, T, variant):
    """The chunk attention kernels (csrc/prefill.cu: tiled k_attn_prefill_t = 2, per-warp k_attn_prefill = 1) against the
    reference's attention() (src/forward.c:141-195) token by token on random q / K / V: causal window 0 .. pos0 + t, ragged
    tiles, every GQA ratio the kernels take, a chunk that starts mid-tile (pos0 % 64 != 0), rows merged from up to five key
    blocks (pos0 = 700). fp32 op tolerance."""
    if variant == 1 and n_heads // n_kv_heads not in (1, 2, 4, 8):
        pytest.skip("the per-warp kernel is instantiated for GQA ratios 1, 2, 4, 8")
    rng = np.random.default_rng(1000 * n_heads + pos0 + T)
    S = pos0 + T
    q = rng.standard_normal((T, n_heads, 128)).astype(np.float32)
    k = rng.standard_normal((S, n_kv_heads * 128)).astype(np.float32)
    v = rng.standard_normal((S, n_kv_heads * 128)).astype(np.float32)
    q[T // 2] *= 3.0  # a peaked softmax row
    got = qlib.attn_prefill(q, k, v, n_heads, n_kv_heads, pos0, variant)
    for t in sorted({0, 1, T // 3, T // 2, T - 2, T - 1} & set(range(T))):
        ref = oracle.attention(q[t].reshape(-1), k, v, n_heads, n_kv_heads, 128, S, pos0 + t)
        close(got[t].reshape(-1), ref, rtol=1e-4, atol=1e-5)


def test_prefill_in_two_calls_equals_one_call(qlib, pkg, ckpt_dir):
    """Prefilling a prompt in two calls (positions 0..a-1, then a..n-1) attends over the first call's cache
    rows and must give the same logits as one call, bit for bit (same kernels, same order per token)."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "tiny-untied", seed=11)
    sh = pkg.checkpoint.SHAPES["tiny-untied"]
    prompt = [int(t) for t in np.random.default_rng(3).integers(0, sh.vocab_size, size=70)]
    with qlib.open(path, 80) as g1, qlib.open(path, 80) as g2:
        a = g1.forward_prefill(prompt, 0)
        g2.forward_prefill(prompt[:33], 0)
        b = g2.forward_prefill(prompt[33:], 33)
        same(a, b)
    # across the internal 512-token chunk boundary: one call (512 + 88) == 200 + 400 == 599 + 1
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "small", seed=11)
    long_prompt = [int(t) for t in np.random.default_rng(4).integers(0, pkg.checkpoint.SHAPES["small"].vocab_size, size=600)]
    with qlib.open(path, 640) as g1, qlib.open(path, 640) as g2, qlib.open(path, 640) as g3:
        a = g1.forward_prefill(long_prompt, 0)
        g2.forward_prefill(long_prompt[:200], 0)
        b = g2.forward_prefill(long_prompt[200:], 200)
        g3.forward_prefill(long_prompt[:599], 0)
        c = g3.forward_prefill(long_prompt[599:], 599)
        same(a, b)
        same(a, c)


def test_persistent_kernel_is_deterministic(qlib, pkg, ckpt_dir):
    """No atomics on data, static work split: the same tokens give bit-identical logits in two independent contexts
    (different arrival orders at every hand-off) and when a context replays the same positions."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "4b-l2", seed=11)
    toks = [3, 1415, 926, 535, 897, 932, 384, 626]
    runs = []
    for _ in range(2):
        with qlib.open(path, 64) as gm:
            a = [gm.forward(t, pos) for pos, t in enumerate(toks)]
            b = [gm.forward(t, pos) for pos, t in enumerate(toks)]  # replay: overwrites the same KV rows with the same values
            for x, y in zip(a, b):
                same(x, y)
            runs.append(a)
    for x, y in zip(*runs):
        same(x, y)


@pytest.mark.parametrize("shape_name", ["4b-l2", "tiny-untied"])
def test_staged_bulk_store_publish_matches_per_warp_stores(qlib, pkg, ckpt_dir, shape_name, monkeypatch):
    """The persistent kernel publishes each CTA's GEMV rows with one TMA bulk store from shared memory (row ranges in
    units of 4 rows); QWEN_MEGA_STAGE=0 selects the older per-warp stores + fence with rows split to the row. Only the way
    results travel differs -- and, since the row ranges differ, which row of a two-row work unit a given row is (the two
    rows of a unit put their tail groups on different lanes, so the fp32 order of a row's group terms can differ in the
    last bit). Logits must agree to fp32 rounding noise, each mode must be bit-reproducible, and a context may switch
    modes between steps (the variable is read per launch)."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, shape_name, seed=11)
    V = pkg.checkpoint.SHAPES[shape_name].vocab_size
    toks = [int(t) for t in np.random.default_rng(8).integers(0, V, size=10)]
    outs = {}
    for mode in ("1", "0", "1b", "0b"):
        monkeypatch.setenv("QWEN_MEGA_STAGE", mode[0])
        with qlib.open(path, 32) as gm:
            outs[mode] = [gm.forward(t, pos) for pos, t in enumerate(toks)]
    for x, y in zip(outs["1"], outs["1b"]):
        same(x, y)
    for x, y in zip(outs["0"], outs["0b"]):
        same(x, y)
    for x, y in zip(outs["1"], outs["0"]):
        assert np.abs(x - y).max() <= NOISE_CAP * max(1.0, float(x.std())), np.abs(x - y).max()
    within = sum(not (np.abs(x - y) > ATOL + RTOL * np.abs(y)).any() for x, y in zip(outs["1"], outs["0"]))
    assert within >= len(toks) - 2, within  # a re-quantisation flip may separate a step or two
    with qlib.open(path, 32) as gm:  # alternate per step: every step must stay close to the single-mode runs
        for pos, t in enumerate(toks):
            monkeypatch.setenv("QWEN_MEGA_STAGE", str(pos & 1))
            lg = gm.forward(t, pos)
            assert np.abs(lg - outs["1"][pos]).max() <= NOISE_CAP * max(1.0, float(lg.std()))


def test_4b_full_shape_first_tokens(qlib, oracle, pkg, ckpt_dir):
    """The headline configuration's checkpoint (Qwen3-4B shape, 36 layers, vocabulary 151936, fast-mode weights as in
    bench.py): the first decode steps of the persistent kernel and a 3-token forward_prefill against the oracle."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "4b", seed=1234, mode="fast")
    toks = [7, 151935, 4711]
    with qlib.open(path, 16) as gm, oracle.open(path, 16) as om:
        ref = [om.forward(t, pos) for pos, t in enumerate(toks)]
        for pos, t in enumerate(toks):
            lg = gm.forward(t, pos)
            assert np.abs(lg - ref[pos]).max() < 0.05 * max(1.0, ref[pos].std()), (pos, np.abs(lg - ref[pos]).max())
            nxt, margin = oracle.argmax(ref[pos])
            assert int(np.argmax(lg)) == nxt or margin < 2e-2
    with qlib.open(path, 16) as gm:
        lg = gm.forward_prefill(toks, 0)
        assert np.abs(lg - ref[-1]).max() < 0.05 * max(1.0, ref[-1].std())


# ---------------------------------------------------------------- device sampler
@pytest.mark.parametrize("grid_kernel", ["1", "0"])
def test_device_sampler_matches_reference_sampler(qlib, oracle, grid_kernel, monkeypatch):
    """Both kernels (QWEN_SAMPLE_GRID=1: cooperative, one CTA per SM; 0: single CTA).
    qwen_cuda_sample_host (csrc/sampler.cu) against the oracle's restatement of sample() (reference src/sampler.c:186-201,
    pinned to the compiled reference in tests/test_sampler_oracle.py) on seeded logits: same token for the same coin.
    The device probabilities differ from the host's in the last bits (expf, summation order), so a token may differ only
    when the oracle reports that the decision sat within 1e-5 (relative) of a boundary -- and that must be rare."""
    monkeypatch.setenv("QWEN_SAMPLE_GRID", grid_kernel)
    rng = np.random.default_rng(2024)
    total = near = 0
    for V in (512, 4096, 151936):
        for t, p in ((1.0, 0.9), (0.7, 0.8), (1e-6, 0.9), (1.5, 0.95), (0.3, 0.5), (1.0, 1e-6)):
            ct, cp = oracle.sampler_clamp(t, p)
            for rep in range(6 if V < 100000 else 3):
                lg = (rng.standard_normal(V) * float(rng.choice([2.0, 6.0, 12.0]))).astype(np.float32)
                coin = float(np.float32(rng.random()))
                want, gap = oracle.sample(lg, ct, cp, coin)
                got = qlib.sample_host(lg, ct, cp, coin)
                if got is None:  # declined: only when the nucleus itself can exceed 4096 tokens (big vocabulary, flat distribution)
                    assert V > 4096 and float(lg.std()) / ct < 4.5, (V, t, p)
                    continue
                total += 1
                if got != want:
                    near += 1
                    assert gap < 1e-5, (V, t, p, coin, got, want, gap)
    assert total > 60 and near <= 2, (total, near)


def test_device_sampler_declines_flat_and_top_p_one(qlib, oracle):
    """More than 4096 tokens can lie in the nucleus: the kernel reports 'not handled' (the caller then runs the
    reference's sample() on host logits) instead of guessing."""
    V = 20000
    flat = np.zeros(V, np.float32)
    assert qlib.sample_host(flat, 1.0, 0.9, 0.5) is None
    lg = (np.random.default_rng(1).standard_normal(V) * 4).astype(np.float32)
    assert qlib.sample_host(lg, 1.0, 1.0, 0.5) is None
    assert qlib.sample_host(lg, 1.0, 0.9, 0.5) == oracle.sample(lg, 1.0, 0.9, 0.5)[0]
    one_hot = np.full(V, -30.0, np.float32)
    one_hot[1234] = 30.0
    for coin in (0.0, 0.999):
        assert qlib.sample_host(one_hot, 1.0, 0.9, coin) == 1234


def test_device_sampler_on_context_logits(qlib, oracle, pkg, ckpt_dir):
    """qwen_cuda_sample on the logits of the last forward(): same tokens as the oracle sampler fed with the host copy of
    those logits and the same xorshift coins; the sampled token drives the next step (a short sampled generation)."""
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "small", seed=11)
    coins, _ = oracle.xorshift_floats(99, 12)
    ct, cp = oracle.sampler_clamp(0.8, 0.9)
    with qlib.open(path, 32) as gm:
        tok = 5
        for pos in range(12):
            lg = gm.forward(tok, pos)
            want, gap = oracle.sample(lg, ct, cp, coins[pos])
            got = gm.sample(ct, cp, coins[pos])
            assert got is not None and (got == want or gap < 1e-5), (pos, got, want, gap)
            tok = want


# ---------------------------------------------------------------- the reference's own CLI on our library (config 1)
def _oracle_completion(oracle, path, ids, n_ctx, temperature, top_p, seed):
    """The reference's completion loop (src/completion.c:57-87) on the oracle: teacher-forced prompt, then sample();
    returns the tokens it prints and how close the run came to a decision boundary (smallest top-2 logit margin and
    smallest relative distance of a coin from a cumulative-mass boundary)."""
    ct, cp = oracle.sampler_clamp(temperature, top_p)
    coins, _ = oracle.xorshift_floats(seed, n_ctx)
    toks, min_margin, min_gap, ci = [], 1e9, 1e9, 0
    with oracle.open(path, n_ctx) as om:
        tok = ids[0]
        for pos in range(n_ctx):
            lo = om.forward(tok, pos)
            toks.append(tok)
            if pos + 1 < len(ids):
                nxt = ids[pos + 1]
            else:
                nxt, gap = oracle.sample(lo, ct, cp, coins[ci])
                ci += 1
                min_margin = min(min_margin, oracle.argmax(lo)[1])
                min_gap = min(min_gap, gap)
            tok = nxt
    return toks, min_margin, min_gap


def test_reference_cli_unmodified_generates_the_same_text(qlib, oracle, pkg, ckpt_dir):
    """BASELINE config 1 in miniature and the drop-in claim of SURVEY.md 8b as a running program: the reference's
    UNMODIFIED examples/qwen.c + src/{qwen,completion,sampler,tokenizer,xorshift}.c, compiled against our headers and
    linked against our libqwen3.so (oracle/_ref/qwen_b200, built by oracle/Makefile where the reference tree exists),
    must print the same completion as the reference's own CPU build (oracle/_ref/qwen_ref): greedy (-t 0) and sampled
    (-t 1 -p 0.9; every logit feeds the host sampler, so the streams match only if the logits do). The prompt / seed
    are chosen so that the oracle's own run stays clear of decision boundaries (asserted), otherwise a one-code
    re-quantisation flip could legitimately change a token."""
    import subprocess
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ours, ref = (os.path.join(here, "oracle", "_ref", n) for n in ("qwen_b200", "qwen_ref"))
    if not (os.path.exists(ours) and os.path.exists(ref)):
        pytest.skip("oracle/_ref/qwen_b200 / qwen_ref not built (they need the reference tree at build time)")
    path = pkg.checkpoint.ensure_checkpoint(ckpt_dir, "tiny-untied", seed=11)
    if not os.path.exists(path + ".tokenizer"):
        pkg.checkpoint.write_tokenizer(path + ".tokenizer", pkg.checkpoint.SHAPES["tiny-untied"].vocab_size)
    for prompt, n_ctx, t, p, seed in (("Qwen", 48, 0.0, 0.9, 1), ("abc", 20, 1.0, 0.9, 33)):
        toks, margin, gap = _oracle_completion(oracle, path, [ord(ch) for ch in prompt], n_ctx, t, p, seed)
        assert (margin > 0.05) if t == 0.0 else (gap > 3e-3), (prompt, margin, gap)  # the chosen run is a fair test
        args = [path, "-m", "completion", "-i", prompt, "-c", str(n_ctx), "-t", str(t), "-p", str(p), "-s", str(seed)]
        outs = []
        for exe in (ref, ours):
            r = subprocess.run([exe] + args, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
            assert r.returncode == 0, (exe, r.stderr[-800:])
            outs.append(r.stdout)
        assert len(set(toks)) > 8, "degenerate completion: not a useful comparison"
        assert outs[0] == outs[1], (prompt, outs[0][:200], outs[1][:200])
