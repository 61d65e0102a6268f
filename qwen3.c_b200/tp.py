"""Tensor-parallel host logic (one process per GPU), SURVEY.md section 8e.

The reference is single process; TP is a new capability that must reproduce the
single-GPU forward. Sharding (identical to what qwen_cuda_create does at upload, csrc/context.cu):

  column-parallel = contiguous ROW blocks of the [out][in] tensors, groups stay whole:
      wq by query head, wk/wv by kv head, w1/w3 and the classifier by rows;
  row-parallel    = COLUMN windows of every row, multiples of 64 so each rank quantises its
      own slice of the attention / SwiGLU output and every int32 group dot is unchanged:
      wo (window of P), w2 (window of hidden_dim);
  one fp32 all-reduce of `dim` floats after wo and after w2; logits all-gathered over vocab rows.
"""
from __future__ import annotations

import ctypes as C
import dataclasses


@dataclasses.dataclass(frozen=True)
class ShardPlan:
    rank: int
    size: int
    q_rows: range      # rows of wq (and slice of the q vector)
    kv_rows: range     # rows of wk / wv
    kv_heads: range
    hid_rows: range    # rows of w1 / w3 == column window of w2
    o_cols: range      # column window of wo
    vocab_rows: range  # rows of the classifier


def shard_plan(shape: dict, rank: int, size: int) -> ShardPlan:
    H, KVH, hd, Hd, V = shape["n_heads"], shape["n_kv_heads"], shape["head_dim"], shape["hidden_dim"], shape["vocab_size"]
    if KVH % size or H % size or V % size or Hd % (64 * size):
        raise ValueError(f"shape not divisible for tp={size}")
    Pl, Kl, Hdl, Vl = H // size * hd, KVH // size * hd, Hd // size, V // size
    return ShardPlan(rank, size, range(rank * Pl, (rank + 1) * Pl), range(rank * Kl, (rank + 1) * Kl),
                     range(rank * KVH // size, (rank + 1) * KVH // size), range(rank * Hdl, (rank + 1) * Hdl),
                     range(rank * Pl, (rank + 1) * Pl), range(rank * Vl, (rank + 1) * Vl))


def _sl(r: range) -> slice:
    return slice(r.start, r.stop)


class EmulatedRank:
    """One rank's decode step written with host ops (numpy + an `ops` object that provides the
    reference's primitives: rmsnorm, q8_quantize, matmul, rotary, swiglu, attention). It exists to
    check the SHARDING MATH on CPU (gloo tests); the GPU path is csrc/decode_ops.cu + NCCL."""

    def __init__(self, views: dict, plan: ShardPlan, ops, allreduce, allgather, seq_len: int, rank_order: bool = False):
        import numpy as np
        self.np, self.v, self.p, self.ops = np, views, plan, ops
        self.allreduce, self.allgather = allreduce, allgather
        # rank_order: the arithmetic of the FUSED device path (csrc/decode_mega.cu, kind 3): every rank receives all
        # partial vectors and adds them to the residual stream one by one in rank order, x = ((x + p0) + p1) + ...,
        # instead of x + allreduce(p) -- bit-identical on every rank by construction
        self.rank_order = rank_order
        L, hd = views["n_layers"], views["head_dim"]
        self.kvd = len(plan.kv_rows)
        self.S = seq_len
        self.k = np.zeros((L, seq_len, self.kvd), np.float32)
        self.vv = np.zeros((L, seq_len, self.kvd), np.float32)
        self.hd = hd

    def _mm(self, x_q, x_s, w, rows: range, cols: range | None = None):
        np = self.np
        q, s = w
        q = q[_sl(rows)]
        s = s[_sl(rows)]
        if cols is not None:
            q = q[:, _sl(cols)]
            s = s[:, cols.start // 64: cols.stop // 64]
        q = np.ascontiguousarray(q)
        s = np.ascontiguousarray(s)
        return self.ops.matmul(x_q, x_s, q.reshape(-1), s.reshape(-1), q.shape[1], q.shape[0])

    def _residual_add(self, x, part):
        if not self.rank_order:
            return x + self.allreduce(part)
        parts = self.allgather(part).reshape(self.p.size, -1)  # slot q = rank q's partial, as in the flow arena
        for q in range(self.p.size):
            x = (x + parts[q]).astype(self.np.float32)
        return x

    def forward(self, token: int, pos: int):
        np, v, p, ops, hd = self.np, self.v, self.p, self.ops, self.hd
        D = v["dim"]
        eq, es = v["emb"]
        x = ops.q8_dequantize(np.ascontiguousarray(eq[token]), np.ascontiguousarray(es[token]))
        n_local_heads = len(p.q_rows) // hd
        n_local_kv = len(p.kv_heads)
        for l in range(v["n_layers"]):
            xq, xs = ops.q8_quantize(ops.rmsnorm(x, v["att_norm"][l]))
            q = self._mm(xq, xs, v["wq"][l], p.q_rows)
            k = self._mm(xq, xs, v["wk"][l], p.kv_rows)
            val = self._mm(xq, xs, v["wv"][l], p.kv_rows)
            for h in range(n_local_heads):
                q[h * hd:(h + 1) * hd] = ops.rotary(ops.rmsnorm(q[h * hd:(h + 1) * hd], v["q_norm"][l]), hd, pos)
            for h in range(n_local_kv):
                k[h * hd:(h + 1) * hd] = ops.rotary(ops.rmsnorm(k[h * hd:(h + 1) * hd], v["k_norm"][l]), hd, pos)
            self.k[l, pos], self.vv[l, pos] = k, val
            att = ops.attention(q, self.k[l], self.vv[l], n_local_heads, n_local_kv, hd, self.S, pos)
            aq, as_ = ops.q8_quantize(att)
            part = self._mm(aq, as_, v["wo"][l], range(0, D), p.o_cols)
            x = self._residual_add(x, part)
            xq, xs = ops.q8_quantize(ops.rmsnorm(x, v["ffn_norm"][l]))
            h1 = self._mm(xq, xs, v["w1"][l], p.hid_rows)
            h3 = self._mm(xq, xs, v["w3"][l], p.hid_rows)
            hq, hs = ops.q8_quantize(ops.swiglu(h1, h3))
            part = self._mm(hq, hs, v["w2"][l], range(0, D), p.hid_rows)
            x = self._residual_add(x, part)
        xq, xs = ops.q8_quantize(ops.rmsnorm(x, v["out_norm"]))
        return self.allgather(self._mm(xq, xs, v["cls"], p.vocab_rows))


def init_tensor_parallel(qlib, model, rank: int, world: int, dist) -> None:
    """Give every rank's context the same NCCL id (rank 0 creates it; `dist` is torch.distributed)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    L = qlib.lib
    L.qwen_cuda_tp_unique_id.argtypes = [C.c_void_p]
    L.qwen_cuda_tp_init.argtypes = [C.c_void_p, C.c_void_p]
    if rank == 0:
        raw = (C.c_ubyte * 128)()
        qlib._ok(L.qwen_cuda_tp_unique_id(raw), "tp_unique_id")
        buf = torch.tensor(list(raw), dtype=torch.uint8)
    if dist.get_backend() == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = buf.to(dev)
        dist.broadcast(t, src=0)
        buf = t.cpu()
    else:
        dist.broadcast(buf, src=0)
    raw = (C.c_ubyte * 128)(*buf.tolist())
    qlib._ok(L.qwen_cuda_tp_init(model.ctx, raw), "tp_init")
