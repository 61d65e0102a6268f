"""Synthetic Q8_0 checkpoints in qwen3.c's on-disk format.

Writes `.bin` files the loader (reference: src/model.c:59-244) accepts, laid out
exactly as the reference exporter does (reference: qwen3/weights.py:249-383):

    [0,48)    12 x int32 header        (weights.py:270-289, include/model.h:30-43)
    [48,256)  zero padding             (weights.py:291-293)
    fp32      att_norm[L][D], ffn_norm[L][D], out_norm[D], q_norm[L][hd], k_norm[L][hd]
    Q8        emb[V][D], wq[l], wk[l], wv[l], wo[l], w1[l], w2[l], w3[l], [cls[V][D]]
    each Q8 tensor = int8[numel] followed by float32[numel/G]   (weights.py:345-347)

and the `.tokenizer` sidecar (reference: qwen3/tokenizer.py:247-278,
src/tokenizer.c:44-109) so the unchanged CLI can open the model.

There is no network, so every benchmark and parity checkpoint is random-init.
Two generators:
  * "gauss": fp32 Gaussian weights pushed through the exporter's weight-side
    quantiser (scale = absmax/127, round-half-even; weights.py:153-160). Used for
    parity: logits get a healthy spread so greedy tokens are well separated.
  * "fast": int8 codes uniform in [-127,127] and scales sigma/73.3*U(0.5,1.5)
    drawn directly. Same bytes-per-token, used for the
    multi-GB shapes where quantising real floats would take minutes.
"""
from __future__ import annotations

import dataclasses
import os
import struct

import numpy as np

MAGIC = 0x7177656E
VERSION = 1
TOK_MAGIC = 0x71746B6E
TOK_VERSION = 2


@dataclasses.dataclass(frozen=True)
class Shape:
    name: str
    dim: int
    hidden_dim: int
    n_layers: int
    n_heads: int
    n_kv_heads: int
    vocab_size: int
    seq_len: int = 40960
    head_dim: int = 128
    shared_classifier: int = 1
    group_size: int = 64

    @property
    def proj_dim(self) -> int:
        return self.n_heads * self.head_dim

    @property
    def kv_dim(self) -> int:
        return self.n_kv_heads * self.head_dim

    def weight_elements(self) -> int:
        """int8 elements read per decoded token (classifier included once)."""
        D, P, K, Hd, L, V = self.dim, self.proj_dim, self.kv_dim, self.hidden_dim, self.n_layers, self.vocab_size
        return L * (2 * D * P + 2 * D * K + 3 * D * Hd) + V * D

    def decode_bytes(self, pos: int, kv_elem_bytes: int = 4) -> int:
        """Algorithmic HBM bytes per decoded token at 0-based position `pos`
        (SURVEY.md section 8d / BASELINE.md section 3)."""
        D, K, L = self.dim, self.kv_dim, self.n_layers
        w = self.weight_elements() * 17 // 16  # int8 + one fp32 scale per 64
        norms = 4 * (2 * L * D + D + 2 * L * self.head_dim)
        kv_read = 2 * L * (pos + 1) * K * kv_elem_bytes
        kv_write = 2 * L * K * kv_elem_bytes
        return w + norms + kv_read + kv_write

    def file_bytes(self) -> int:
        D, P, K, Hd, L, V, G = (self.dim, self.proj_dim, self.kv_dim, self.hidden_dim, self.n_layers,
                                self.vocab_size, self.group_size)
        q8 = V * D + L * (2 * D * P + 2 * D * K + 3 * D * Hd) + (0 if self.shared_classifier else V * D)
        return 256 + 4 * (2 * L * D + D + 2 * L * self.head_dim) + q8 + 4 * (q8 // G)


# Shapes of BASELINE.json's five configs ("[recalled]") + tiny test shapes.
SHAPES = {
    "tiny": Shape("tiny", 256, 768, 2, 4, 2, 512, seq_len=256),
    "tiny-untied": Shape("tiny-untied", 256, 512, 3, 8, 2, 384, seq_len=192, shared_classifier=0),
    "small": Shape("small", 512, 1536, 4, 8, 4, 2048, seq_len=1024),
    "0.6b-l1": Shape("0.6b-l1", 1024, 3072, 1, 16, 8, 4096, seq_len=256),
    # two layers of the real 4B / 8B / 32B layer shapes (GQA 4, 4, 8; hidden 9728 / 12288 / 25600) with a small vocabulary
    "4b-l2": Shape("4b-l2", 2560, 9728, 2, 32, 8, 4096, seq_len=4096),
    "8b-l2": Shape("8b-l2", 4096, 12288, 2, 32, 8, 4096, seq_len=4096, shared_classifier=0),
    "32b-l2": Shape("32b-l2", 5120, 25600, 2, 64, 8, 4096, seq_len=4096, shared_classifier=0),
    "0.6b": Shape("0.6b", 1024, 3072, 28, 16, 8, 151936),
    "1.7b": Shape("1.7b", 2048, 6144, 28, 16, 8, 151936),
    "4b": Shape("4b", 2560, 9728, 36, 32, 8, 151936),
    "8b": Shape("8b", 4096, 12288, 36, 32, 8, 151936, shared_classifier=0),
    "32b": Shape("32b", 5120, 25600, 64, 64, 8, 151936, shared_classifier=0),
}


def quantize_q8_0(w: np.ndarray, group: int) -> tuple[np.ndarray, np.ndarray]:
    """Weight-side quantiser, numpy restatement of reference qwen3/weights.py:137-166
    (np.round is round-half-even like torch.round)."""
    g = w.astype(np.float32).reshape(-1, group)
    amax = np.abs(g).max(axis=1)
    scale = (amax / np.float32(127.0)).astype(np.float32)
    q = np.round(g / scale[:, None]).astype(np.int8)
    return q.reshape(-1), scale


def _q8_gauss(rng: np.random.Generator, numel: int, group: int, sigma: float):
    # generate in slabs so a 400 MB tensor does not need 3 GB of temporaries
    slab = 1 << 24
    qs, ss = [], []
    for start in range(0, numel, slab):
        n = min(slab, numel - start)
        w = rng.standard_normal(n, dtype=np.float32) * np.float32(sigma)
        q, s = quantize_q8_0(w, group)
        qs.append(q)
        ss.append(s)
    return np.concatenate(qs), np.concatenate(ss)


def _q8_fast(rng: np.random.Generator, numel: int, group: int, sigma: float):
    # 8 codes per 64-bit draw (~8x faster than int8 draws); -128 is folded onto -127 (Q8_0 never emits it)
    i64 = np.iinfo(np.int64)
    q = rng.integers(i64.min, i64.max, size=(numel + 7) // 8, dtype=np.int64, endpoint=True).view(np.int8)[:numel]
    np.maximum(q, -127, out=q)
    s = (np.float32(sigma / 73.3) * rng.uniform(0.5, 1.5, size=numel // group)).astype(np.float32)
    return q, s


def write_checkpoint(path: str, shape: Shape, seed: int = 1234, mode: str = "gauss",
                     sigma: float = 0.02, emb_sigma: float | None = None, rho: float = 0.03) -> str:
    """Write a random-init checkpoint; returns `path`. Deterministic in (shape, seed, mode, rho).

    Conditioning (SURVEY.md H7 and DESIGN.md "synthetic checkpoints"): a random-init residual
    network whose branches are as large as its residual stream doubles any perturbation per
    block, so one flipped int8 activation code (which the reference's own -O2 and -Ofast builds
    already disagree on) grows to O(0.1) logit differences after 28+ layers. Real checkpoints
    are not like that: branch outputs are small against the stream. The generator therefore
    sizes the output projections (wo, w2) so each branch adds about `rho` x the stream's rms,
    and the embedding so logits have std ~6 (greedy tokens well separated)."""
    rng = np.random.default_rng(seed)
    gen = _q8_gauss if mode == "gauss" else _q8_fast
    D, P, K, Hd, L, V, G, hd = (shape.dim, shape.proj_dim, shape.kv_dim, shape.hidden_dim, shape.n_layers,
                                shape.vocab_size, shape.group_size, shape.head_dim)
    emb_sigma = (6.0 / float(np.sqrt(D))) if emb_sigma is None else emb_sigma
    # branch gains: attention output ~ rms(v) ~ sigma*sqrt(D); swiglu output rms ~ 0.3*(sigma*sqrt(D))
    v_rms = sigma * float(np.sqrt(D))
    wo_sigma = rho * emb_sigma / (float(np.sqrt(P)) * v_rms)
    w2_sigma = rho * emb_sigma / (float(np.sqrt(Hd)) * 0.3 * v_rms)
    tmp = path + ".part"
    with open(tmp, "wb") as f:
        f.write(struct.pack("<12i", MAGIC, VERSION, D, Hd, L, shape.n_heads, shape.n_kv_heads, V,
                            shape.seq_len, hd, shape.shared_classifier, G))
        f.write(b"\0" * (256 - 48))

        def norm(n):
            return (1.0 + 0.05 * rng.standard_normal(n)).astype(np.float32)

        f.write(norm(L * D).tobytes())    # att
        f.write(norm(L * D).tobytes())    # ffn
        f.write(norm(D).tobytes())        # out
        f.write(norm(L * hd).tobytes())   # q
        f.write(norm(L * hd).tobytes())   # k

        def q8(numel, sg):
            q, s = gen(rng, numel, G, sg)
            f.write(q.tobytes())
            f.write(s.tobytes())

        q8(V * D, emb_sigma)
        for numel, sg in ((D * P, sigma), (D * K, sigma), (D * K, sigma), (P * D, wo_sigma), (D * Hd, sigma),
                          (Hd * D, w2_sigma), (D * Hd, sigma)):
            for _ in range(L):
                q8(numel, sg)
        if not shape.shared_classifier:
            q8(V * D, emb_sigma)
    os.replace(tmp, path)
    assert os.path.getsize(path) == shape.file_bytes(), (os.path.getsize(path), shape.file_bytes())
    return path


def write_tokenizer(path: str, vocab_size: int) -> str:
    """Synthetic `.tokenizer`: every id decodes to a short ASCII string, every special id is -1
    so generation never stops early (reference: src/tokenizer.c:44-109, include/tokenizer.h:45-60)."""
    with open(path, "wb") as f:
        max_len = 8
        f.write(struct.pack("<Iiii", TOK_MAGIC, TOK_VERSION, vocab_size, max_len))
        f.write(struct.pack("<10i", *([-1] * 10)))
        for i in range(vocab_size):
            tok = (b"t%x " % i)[:max_len] if i >= 256 else bytes([i]) if 32 <= i < 127 else b"<%02x>" % i
            f.write(struct.pack("<fi", -float(np.log(i + 1.0)), len(tok)))
            f.write(tok)
    return path


def ensure_checkpoint(directory: str, shape_name: str, seed: int = 1234, mode: str = "gauss", **kw) -> str:
    """Create `<directory>/<shape>-<mode>-<seed>.bin` if missing and return its path."""
    os.makedirs(directory, exist_ok=True)
    shape = SHAPES[shape_name]
    path = os.path.join(directory, f"qwen3-{shape.name}-{mode}-{seed}.bin")
    if not (os.path.exists(path) and os.path.getsize(path) == shape.file_bytes()):
        write_checkpoint(path, shape, seed=seed, mode=mode, **kw)
    return path


def load_views(path: str) -> dict:
    """Memory-map a `.bin` and return its header fields plus numpy views of every tensor in export
    order (reference reader: src/model.c:59-244). Used by the tensor-parallel host logic and tests."""
    hdr = np.fromfile(path, dtype="<i4", count=12)
    assert hdr[0] == MAGIC and hdr[1] == VERSION, "not a qwen3.c checkpoint"
    D, Hd, L, H, KVH, V, S, hd, shared, G = (int(x) for x in hdr[2:12])
    P, K = H * hd, KVH * hd
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    off = 256
    out = dict(dim=D, hidden_dim=Hd, n_layers=L, n_heads=H, n_kv_heads=KVH, vocab_size=V, seq_len=S, head_dim=hd,
               shared_classifier=shared, group_size=G)

    def f32(n):
        nonlocal off
        a = mm[off: off + 4 * n].view(np.float32)
        off += 4 * n
        return a

    def q8(rows, cols):
        nonlocal off
        n = rows * cols
        q = mm[off: off + n].view(np.int8).reshape(rows, cols)
        off += n
        s = mm[off: off + 4 * (n // G)].view(np.float32).reshape(rows, cols // G)
        off += 4 * (n // G)
        return q, s

    out["att_norm"] = f32(L * D).reshape(L, D)
    out["ffn_norm"] = f32(L * D).reshape(L, D)
    out["out_norm"] = f32(D)
    out["q_norm"] = f32(L * hd).reshape(L, hd)
    out["k_norm"] = f32(L * hd).reshape(L, hd)
    out["emb"] = q8(V, D)
    for name, rows, cols in (("wq", P, D), ("wk", K, D), ("wv", K, D), ("wo", D, P), ("w1", Hd, D), ("w2", D, Hd),
                             ("w3", Hd, D)):
        out[name] = [q8(rows, cols) for _ in range(L)]
    out["cls"] = out["emb"] if shared else q8(V, D)
    assert off == mm.size, (off, mm.size)
    return out
