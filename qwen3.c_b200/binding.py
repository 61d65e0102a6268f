"""ctypes mirror of the reference's operator interface, bound to this repo's libqwen3.so.

Function names, argument order and meaning follow include/forward.h, include/q8.h and
include/model.h (reference: include/forward.h:31-140, q8.h:25-30, model.h:153-168), so a
parity test written against the reference library reads the same against this one. The
qwen_cuda_* shim (include/qwen_cuda.h) is bound as well for device-resident runs.

No computation happens in this module and nothing here falls back to the CPU: if the
library is missing or no CUDA device is visible, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import build as _build

c_float_p = C.POINTER(C.c_float)
c_int8_p = C.POINTER(C.c_int8)
c_int32_p = C.POINTER(C.c_int32)


class Q8Tensor(C.Structure):  # include/q8.h
    _fields_ = [("s", c_float_p), ("q", c_int8_p)]


class ModelParams(C.Structure):  # include/model.h
    _fields_ = [(n, C.c_int) for n in ("magic", "version", "dim", "hidden_dim", "n_layers", "n_heads",
                                       "n_kv_heads", "vocab_size", "seq_len", "head_dim",
                                       "shared_classifier", "block_size")]


class ModelWeights(C.Structure):
    _fields_ = ([(n, C.POINTER(Q8Tensor)) for n in ("wq", "wk", "wv", "wo", "w1", "w2", "w3", "cls", "qe")]
                + [(n, c_float_p) for n in ("fe", "att_rms_norm", "ffn_rms_norm", "out_rms_norm",
                                            "q_rms_norm", "k_rms_norm")])


class ForwardState(C.Structure):
    _fields_ = ([(n, c_float_p) for n in ("x", "x_rms_norm", "q", "k", "v", "scores", "mlp_in", "mlp_gate",
                                          "logits", "k_cache", "v_cache")]
                + [("qx", Q8Tensor), ("qh", Q8Tensor)])


class Model(C.Structure):
    _fields_ = [("params", ModelParams), ("weights", ModelWeights), ("state", ForwardState),
                ("data", C.c_void_p), ("size", C.c_ssize_t)]


def _fp(a):
    assert a.dtype == np.float32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_float_p)


def _i8(a):
    assert a.dtype == np.int8 and a.flags.c_contiguous
    return a.ctypes.data_as(c_int8_p)


class QwenLib:
    """libqwen3.so with prototypes attached."""

    def __init__(self, path: str | None = None):
        self.path = path or _build.lib_path()
        L = self.lib = C.CDLL(self.path)
        L.model_create.restype = C.POINTER(Model)
        L.model_create.argtypes = [C.c_char_p, C.c_int]
        L.model_free.argtypes = [C.POINTER(Model)]
        L.model_free.restype = None
        L.model_cuda_ctx.restype = C.c_void_p
        L.model_cuda_ctx.argtypes = [C.POINTER(Model)]
        L.forward.restype = c_float_p
        L.forward.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
        L.forward_prefill.restype = c_float_p
        L.forward_prefill.argtypes = [C.POINTER(Model), C.POINTER(C.c_int), C.c_int, C.c_int]
        L.qwen_cuda_prefill.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.c_int, c_float_p]
        L.attention.argtypes = [C.POINTER(Model), C.c_int, C.c_int]
        L.attention.restype = None
        L.rmsnorm.argtypes = [c_float_p, c_float_p, c_float_p, C.c_int]
        L.softmax.argtypes = [c_float_p, C.c_int]
        L.matmul.argtypes = [c_float_p, C.POINTER(Q8Tensor), C.POINTER(Q8Tensor), C.c_int, C.c_int, C.c_int]
        L.rotary.argtypes = [c_float_p, C.c_int, C.c_int]
        L.swiglu.argtypes = [c_float_p, c_float_p, C.c_int]
        for f in (L.rmsnorm, L.softmax, L.matmul, L.rotary, L.swiglu):
            f.restype = None
        L.sigmoid.argtypes = [C.c_float]
        L.sigmoid.restype = C.c_float
        L.silu.argtypes = [C.c_float]
        L.silu.restype = C.c_float
        L.q8_quantize.argtypes = [C.POINTER(Q8Tensor), c_float_p, C.c_int, C.c_int]
        L.q8_quantize.restype = None
        L.q8_dequantize.argtypes = [C.POINTER(Q8Tensor), c_float_p, C.c_int, C.c_int]
        L.q8_dequantize.restype = None
        # shim
        L.qwen_cuda_last_error.restype = C.c_char_p
        L.qwen_cuda_device_count.restype = C.c_int
        L.qwen_cuda_forward.argtypes = [C.c_void_p, C.c_int, C.c_int, c_float_p]
        L.qwen_cuda_forward_async.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.qwen_cuda_logits_to_host.argtypes = [C.c_void_p, c_float_p]
        L.qwen_cuda_decode_greedy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_int32_p]
        L.qwen_cuda_time_decode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, c_float_p, c_int32_p]
        L.qwen_cuda_sync.argtypes = [C.c_void_p]
        L.qwen_cuda_set_path.argtypes = [C.c_void_p, C.c_int]
        L.qwen_cuda_get_path.argtypes = [C.c_void_p]
        L.qwen_cuda_sample.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, c_int32_p]
        L.qwen_cuda_sample_host.argtypes = [c_float_p, C.c_int, C.c_float, C.c_float, C.c_float, c_int32_p]
        L.qwen_cuda_kv_write.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p]
        L.qwen_cuda_kv_read.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, c_float_p, c_float_p]
        L.qwen_cuda_debug_set_layers.argtypes = [C.c_void_p, C.c_int]
        L.qwen_cuda_debug_set_window.argtypes = [C.c_void_p, C.c_int, C.c_int, c_float_p]
        L.qwen_cuda_debug_codes_enable.argtypes = [C.c_void_p, C.c_int]
        L.qwen_cuda_debug_codes_read.argtypes = [C.c_void_p, C.c_int, c_int8_p, c_float_p, C.c_int]
        L.qwen_cuda_debug_quantize_fused.argtypes = [c_int8_p, c_float_p, c_float_p, C.c_int]
        L.qwen_cuda_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t]
        L.qwen_cuda_matmul_group_dots.argtypes = [c_int32_p, c_int8_p, c_int8_p, C.c_int, C.c_int, C.c_int]
        L.qwen_cuda_attention.argtypes = [C.c_void_p, C.c_int, C.c_int, c_float_p, c_float_p]
        L.qwen_cuda_debug_attn_prefill.argtypes = [c_float_p, c_float_p, c_float_p, c_float_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.qwen_cuda_matmul_batch.argtypes = [c_float_p, c_int32_p, c_int8_p, c_float_p, c_int8_p, c_float_p, C.c_int, C.c_int,
                                             C.c_int, C.c_int, c_float_p]

    def sample_host(self, logits, temperature: float, top_p: float, coin: float):
        """Device sampler on host logits (test hook): returns the token, or None when the fast path does not apply."""
        x = np.ascontiguousarray(logits, np.float32)
        tok = C.c_int32(-1)
        rc = self.lib.qwen_cuda_sample_host(_fp(x), x.size, temperature, top_p, coin, C.byref(tok))
        if rc < 0:
            raise RuntimeError("sample_host: " + self.err())
        return int(tok.value) if rc == 0 else None

    def err(self) -> str:
        return (self.lib.qwen_cuda_last_error() or b"").decode()

    def _void(self, fn, *args):
        """Call a void reference-API function; raise if the device layer reported a failure."""
        self.lib.qwen_cuda_clear_error()
        fn(*args)
        e = self.err()
        if e:
            raise RuntimeError(e)

    def _ok(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.err()}")

    # ---- ops: reference names, numpy in/out -----------------------------------
    def q8_quantize(self, x, gs=64):
        x = np.ascontiguousarray(x, np.float32)
        q = np.zeros(x.size, np.int8)
        s = np.zeros(max(x.size // gs, 1), np.float32)
        t = Q8Tensor(_fp(s), _i8(q))
        self._void(self.lib.q8_quantize, C.byref(t), _fp(x), x.size, gs)
        return q, s[: x.size // gs]

    def q8_dequantize(self, q, s, gs=64):
        x = np.zeros(q.size, np.float32)
        t = Q8Tensor(_fp(s), _i8(q))
        self._void(self.lib.q8_dequantize, C.byref(t), _fp(x), q.size, gs)
        return x

    def matmul(self, xq, xs, wq, ws, n, d, gs=64):
        out = np.full(d, np.nan, np.float32)
        tx, tw = Q8Tensor(_fp(xs), _i8(xq)), Q8Tensor(_fp(ws), _i8(wq))
        self._void(self.lib.matmul, _fp(out), C.byref(tx), C.byref(tw), n, d, gs)
        return out

    def group_dots(self, xq, wq, n, d, gs=64):
        dots = np.zeros(d * (n // gs), np.int32)
        self._ok(self.lib.qwen_cuda_matmul_group_dots(dots.ctypes.data_as(c_int32_p), _i8(xq), _i8(wq), n, d, gs),
                 "group_dots")
        return dots.reshape(d, n // gs)

    def matmul_batch(self, xq, xs, wq, ws, n, d, T, want_dots=False, reps=1):
        """T-token matmul on the tcgen05 tensor cores (prefill). Returns (out[T][d], dots or None, best ms)."""
        out = np.full((T, d), np.nan, np.float32)
        dots = np.zeros((T, d, n // 64), np.int32) if want_dots else None
        ms = C.c_float(0)
        self._ok(self.lib.qwen_cuda_matmul_batch(_fp(out), dots.ctypes.data_as(c_int32_p) if want_dots else None,
                                                 _i8(np.ascontiguousarray(xq)), _fp(np.ascontiguousarray(xs)), _i8(wq), _fp(ws),
                                                 n, d, T, reps, C.byref(ms)), "matmul_batch")
        return out, dots, ms.value

    def attn_prefill(self, q, k, v, n_heads, n_kv_heads, pos0, variant=2):
        """Chunk attention kernel on host data (test hook): q [T][n_heads][128], k / v [pos0 + T][n_kv_heads * 128]."""
        q = np.ascontiguousarray(q, np.float32)
        T = q.shape[0]
        out = np.full(q.shape, np.nan, np.float32)
        self._ok(self.lib.qwen_cuda_debug_attn_prefill(_fp(out), _fp(q), _fp(np.ascontiguousarray(k, np.float32)),
                                                       _fp(np.ascontiguousarray(v, np.float32)), n_heads, n_kv_heads, pos0, T, variant),
                 "debug_attn_prefill")
        return out

    def quantize_fused(self, x):
        """The persistent decode kernel's fused quantiser on a host vector (test hook)."""
        x = np.ascontiguousarray(x, np.float32)
        q = np.zeros(x.size, np.int8)
        s = np.zeros(x.size // 64, np.float32)
        self._ok(self.lib.qwen_cuda_debug_quantize_fused(_i8(q), _fp(s), _fp(x), x.size), "quantize_fused")
        return q, s

    def rmsnorm(self, x, w):
        x = np.ascontiguousarray(x, np.float32)
        out = np.full(x.size, np.nan, np.float32)
        self._void(self.lib.rmsnorm, _fp(out), _fp(x), _fp(np.ascontiguousarray(w, np.float32)), x.size)
        return out

    def softmax(self, x):
        y = np.array(x, np.float32, copy=True)
        self._void(self.lib.softmax, _fp(y), y.size)
        return y

    def rotary(self, x, head_dim, pos):
        y = np.array(x, np.float32, copy=True)
        self._void(self.lib.rotary, _fp(y), head_dim, pos)
        return y

    def swiglu(self, x1, x3):
        y = np.array(x1, np.float32, copy=True)
        self._void(self.lib.swiglu, _fp(y), _fp(np.ascontiguousarray(x3, np.float32)), y.size)
        return y

    def open(self, path: str, seq_len: int = 0) -> "B200Model":
        return B200Model(self, path, seq_len)


class B200Model:
    """A checkpoint resident on the B200, driven through model_create / forward / model_free."""

    def __init__(self, ql: QwenLib, path: str, seq_len: int = 0):
        self.ql = ql
        self.m = ql.lib.model_create(path.encode(), seq_len)
        if not self.m:
            raise RuntimeError(f"model_create({path}) failed: {ql.err()}")
        self.p = self.m.contents.params
        self.ctx = ql.lib.model_cuda_ctx(self.m)

    def forward(self, token: int, pos: int) -> np.ndarray:
        ptr = self.ql.lib.forward(self.m, token, pos)
        if not ptr:
            raise RuntimeError(f"forward failed: {self.ql.err()}")
        return np.ctypeslib.as_array(ptr, shape=(self.p.vocab_size,)).copy()

    def forward_prefill(self, tokens, pos: int) -> np.ndarray:
        """forward_prefill(): the prompt tokens at positions pos.. in one pass; logits of the last token."""
        arr = (C.c_int * len(tokens))(*[int(t) for t in tokens])
        ptr = self.ql.lib.forward_prefill(self.m, arr, len(tokens), pos)
        if not ptr:
            raise RuntimeError(f"forward_prefill failed: {self.ql.err()}")
        return np.ctypeslib.as_array(ptr, shape=(self.p.vocab_size,)).copy()

    def prefill_nocopy(self, tokens, pos: int) -> bool:
        """Prefill without the logits copy (device-side timing)."""
        arr = (C.c_int * len(tokens))(*[int(t) for t in tokens])
        return self.ql.lib.qwen_cuda_prefill(self.ctx, arr, len(tokens), pos, None) == 0

    def forward_nocopy(self, token: int, pos: int):
        """forward() exactly as a C caller sees it: returns the pinned logits pointer."""
        return self.ql.lib.forward(self.m, token, pos)

    def set_path(self, path: int):
        self.ql._ok(self.ql.lib.qwen_cuda_set_path(self.ctx, path), "set_path")

    def forward_async(self, token: int, pos: int) -> None:
        """Enqueue one decode step; the logits stay on the device (qwen_cuda_forward_async)."""
        self.ql._ok(self.ql.lib.qwen_cuda_forward_async(self.ctx, token, pos), "forward_async")

    def sample(self, temperature: float, top_p: float, coin: float):
        """sample() on the device logits of the last step (reference src/sampler.c:186-201); None = use the host path."""
        tok = C.c_int32(-1)
        rc = self.ql.lib.qwen_cuda_sample(self.ctx, temperature, top_p, coin, C.byref(tok))
        if rc < 0:
            raise RuntimeError("sample: " + self.ql.err())
        return int(tok.value) if rc == 0 else None

    def get_path(self) -> int:
        return int(self.ql.lib.qwen_cuda_get_path(self.ctx))

    def set_layers(self, n: int):
        self.ql._ok(self.ql.lib.qwen_cuda_debug_set_layers(self.ctx, n), "set_layers")

    def set_window(self, l0: int = 0, l1: int = -1, x=None):
        """Debug: the next steps run layers [l0, l1) only; x (dim floats) replaces the embedding row entering layer l0."""
        xp = None
        if x is not None:
            x = np.ascontiguousarray(x, np.float32)
            assert x.size == self.p.dim
            xp = _fp(x)
        self.ql._ok(self.ql.lib.qwen_cuda_debug_set_window(self.ctx, l0, l1, xp), "set_window")

    def codes_enable(self, on: bool = True):
        self.ql._ok(self.ql.lib.qwen_cuda_debug_codes_enable(self.ctx, int(on)), "codes_enable")

    def codes_read(self, which: int, n: int):
        """Q8_0 codes + scales the persistent kernel fed to GEMV `which` (4 * layer + {0 qkv, 1 wo, 2 w1/w3, 3 w2};
        4 * n_layers = classifier) in the last step."""
        q = np.zeros(n, np.int8)
        sc = np.zeros(n // 64, np.float32)
        self.ql._ok(self.ql.lib.qwen_cuda_debug_codes_read(self.ctx, which, q.ctypes.data_as(c_int8_p), _fp(sc), n), "codes_read")
        return q, sc

    def decode_greedy(self, first_token: int, pos0: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.int32)
        self.ql._ok(self.ql.lib.qwen_cuda_decode_greedy(self.ctx, first_token, pos0, n, out.ctypes.data_as(c_int32_p)),
                    "decode_greedy")
        return out

    def time_decode(self, token: int, pos0: int, steps: int, warmup: int):
        ms, launches = C.c_float(0), C.c_int32(0)
        self.ql._ok(self.ql.lib.qwen_cuda_time_decode(self.ctx, token, pos0, steps, warmup, C.byref(ms),
                                                      C.byref(launches)), "time_decode")
        return ms.value, launches.value

    def kv_write(self, layer, pos0, k, v):
        k = np.ascontiguousarray(k, np.float32)
        v = np.ascontiguousarray(v, np.float32)
        self.ql._ok(self.ql.lib.qwen_cuda_kv_write(self.ctx, layer, pos0, k.shape[0], _fp(k), _fp(v)), "kv_write")

    def kv_read(self, layer, pos0, npos):
        kvd = self.p.n_kv_heads * self.p.head_dim
        k = np.zeros((npos, kvd), np.float32)
        v = np.zeros((npos, kvd), np.float32)
        self.ql._ok(self.ql.lib.qwen_cuda_kv_read(self.ctx, layer, pos0, npos, _fp(k), _fp(v)), "kv_read")
        return k, v

    def debug_read(self, what: str, n: int, dtype=np.float32) -> np.ndarray:
        buf = np.zeros(n, dtype)
        rc = self.ql.lib.qwen_cuda_debug_read(self.ctx, what.encode(), buf.ctypes.data_as(C.c_void_p), buf.nbytes)
        if rc < 0:
            raise RuntimeError(self.ql.err())
        return buf

    def attention(self, layer: int, pos: int, q: np.ndarray) -> np.ndarray:
        """attention() as the reference declares it: q in state.q, result in state.x_rms_norm."""
        P = self.p.n_heads * self.p.head_dim
        C.memmove(self.m.contents.state.q, np.ascontiguousarray(q, np.float32).ctypes.data, 4 * P)
        self.ql._void(self.ql.lib.attention, self.m, layer, pos)
        return np.ctypeslib.as_array(self.m.contents.state.x_rms_norm, shape=(P,)).copy()

    def close(self):
        if self.m:
            self.ql.lib.model_free(self.m)
            self.m = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
